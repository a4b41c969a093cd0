"""The C-ABI library loads, exports every symbol include/amsweep.h declares,
agrees with the oracle's struct layouts, and FAILS LOUDLY without a GPU."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "amsweep.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(am_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(am, lib):
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in amsweep.h but not exported"
    # and the Python binding types every one of them
    assert sorted(am.abi.SYMBOLS) == declared


def test_abi_version_and_strerror(am, lib):
    assert lib.am_abi_version() == 2
    assert lib.am_strerror(0) == b"ok"
    for code in range(-8, 0):
        assert lib.am_strerror(code) not in (b"", b"unknown amsweep error")


def test_struct_layouts_match_oracle(am, orc):
    assert C.sizeof(am.AmCron) == C.sizeof(orc.OrcCron) == 56
    assert C.sizeof(am.AmRecord) == 96
    assert C.sizeof(am.AmTickStats) == C.sizeof(orc.OrcStats) == 128
    assert C.sizeof(am.AmRecordCols) == C.sizeof(orc.OrcCols) == 16 * 8
    import amgen
    assert C.sizeof(am.AmHealthCheck) == C.sizeof(amgen.HealthCheckC) == 120  # static_assert-ed in C
    assert [n for n, _ in orc.COLUMNS] == am.COLUMN_NAMES


def test_invalid_arguments_are_codes_not_crashes(am, lib):
    assert lib.am_cron_parse(None, 0, None, None, 0) == am.AM_E_INVAL
    assert lib.am_healthcheck_classify(None, None) == am.AM_E_INVAL
    assert lib.am_sweep_create(None, 0, 10, 0) == am.AM_E_INVAL
    assert lib.am_sweep_size(None) == 0
    assert lib.am_sweep_tick(None, 0, 0, None, None, 0, None, None) == am.AM_E_INVAL


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_create_fails_loudly_without_gpu(am):
    """No CPU fallback: without a device the product path refuses to exist."""
    with pytest.raises(am.AmError) as ei:
        am.Sweep(capacity=1024)
    assert ei.value.code == am.AM_E_DEVICE
    # the reason is readable from ANY thread afterwards (a cgo caller may have been moved
    # to another OS thread between the two calls)
    import threading
    seen = []
    t = threading.Thread(target=lambda: seen.append(am.load().am_last_error_detail(None)))
    t.start()
    t.join()
    assert seen and seen[0] and b"cuda" in seen[0].lower()


def test_product_never_links_or_imports_the_oracle(am):
    """oracle/ is test infrastructure: the shipped .so and package must not reference it."""
    blob = open(am.abi.LIB_PATH, "rb").read()
    assert b"orc_" not in blob and b"amsweep_oracle" not in blob
    pkg = os.path.dirname(am.abi.LIB_PATH)
    pkg = os.path.dirname(pkg)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f == "build.py":
                continue  # compiles the checker (allowed); never loads or calls it
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_c" not in src and "oracle_py" not in src and "amsweep_oracle" not in src, f
