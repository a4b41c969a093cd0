"""Frozen vectors (tests/golden/*.json, made by tests/golden/make_golden.py):
the C oracle must keep reproducing them (CPU), and so must the CUDA sweep (GPU)."""
import glob
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import ROOT

FIXTURES = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.json")))


def _digest(cols, names):
    h = hashlib.sha256()
    for name in names:
        h.update(np.ascontiguousarray(cols[name]).tobytes())
    return h.hexdigest()


def _check_tick(entry, idx, act, st, cols, names, who):
    assert st == entry["stats"], f"{who}: stats at T={entry['T']}"
    if "idx" in entry:
        assert idx.tolist() == entry["idx"] and act.tolist() == entry["act"], f"{who}: list at T={entry['T']}"
    else:
        got = hashlib.sha256(idx.astype(np.uint64).tobytes() + act.astype(np.uint32).tobytes()).hexdigest()
        assert got == entry["idx_act_sha256"], f"{who}: list digest at T={entry['T']}"
    assert _digest(cols, names) == entry["columns_sha256"], f"{who}: columns at T={entry['T']}"


def test_fixtures_exist():
    assert len(FIXTURES) >= 5


@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_oracle_reproduces_golden(orc, gen, path):
    fx = json.load(open(path))
    names = [n for n, _ in orc.COLUMNS]
    cols = gen.fill(fx["config"], fx["seed"], 0, fx["n"], fx["T0"], orc.load().orc_classify)
    assert _digest(cols, names) == fx["initial_columns_sha256"], "generator or classifier drifted"
    for entry in fx["ticks"]:
        idx, act, st = orc.sweep(cols, entry["T"], mode=entry["mode"], seed=fx["seed"])
        _check_tick(entry, idx, act, st, cols, names, "oracle")


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=[os.path.basename(p)[:-5] for p in FIXTURES])
def test_cuda_sweep_reproduces_golden(am, gen, path):
    fx = json.load(open(path))
    cols = gen.fill(fx["config"], fx["seed"], 0, fx["n"], fx["T0"], am.load().am_healthcheck_classify)
    assert _digest(cols, am.COLUMN_NAMES) == fx["initial_columns_sha256"], "product classifier != frozen"
    with am.Sweep(capacity=fx["n"]) as s:
        s.load_range(0, cols)
        s.set_seed(fx["seed"])
        for entry in fx["ticks"]:
            idx, act, st = s.tick(entry["T"], mode=entry["mode"])
            dev = s.read_range(0, fx["n"])
            _check_tick(entry, idx, act, st, dev, am.COLUMN_NAMES, "cuda")
