"""The NVLink exchange kernels (csrc/gather_kernels.cuh) on a CPU emulation of the CUDA execution
model (tests/emu/cuda_emu.h): ranks are OS threads, a CTA's threads are fibers, `__syncthreads`
and the warp collectives are rendezvous points, system-scope acquire/release are std atomics.

The SAME kernel source the GPU runs is compiled for the host; every rank pushes several ticks
back to back (no host barrier, as on the stream) and checks that its output is the rank-ordered
concatenation of all ranks' lists.  This covers the N>1 path's index arithmetic, buffer layout
and epoch / ticket / done-flag protocol without a GPU — for the two formats validated on
hardware (plain, c3: which also validates the emulator) and for the experimental bitmap format
that has not run on hardware yet.  It does not cover the GPU memory model or performance.
"""
import os
import subprocess

import pytest

from conftest import ROOT

EMU = os.path.join(ROOT, "tests", "emu")
EXE = os.path.join(EMU, "emu_gather.bin")


@pytest.fixture(scope="module")
def emu_bin():
    subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-Wno-unknown-pragmas", "-pthread",
                    "-I", os.path.join(ROOT, "include"), os.path.join(EMU, "emu_gather.cpp"), "-o", EXE],
                   check=True)
    return EXE


# world, idx_bytes, records per rank, epochs, CTAs, density %, [capacity override]
CASES = [
    (2, 4, 20000, 4, 3, 33),
    (2, 8, 20000, 6, 5, 33),
    (3, 4, 30011, 5, 4, 33),      # ragged shard sizes, groups straddling nothing: 3 full + 1 partial
    (4, 4, 9000, 5, 7, 50),
    (8, 4, 8192, 4, 2, 33),       # exactly one group per rank
    (8, 8, 5000, 3, 16, 10),      # more CTAs than groups
    (2, 4, 100, 3, 3, 100),       # every record emitted
    (3, 8, 8191, 3, 1, 1),        # one CTA, nearly empty lists
    (2, 4, 16384, 3, 40, 0),      # only the always-dense / always-empty stretches
    (3, 4, 20000, 4, 3, 50, 21000),   # capacity below the total: truncation paths
    (2, 4, 600000, 2, 1, 33),     # 74 groups on ONE CTA (c3: no CTAs to split off for the searches)
]


@pytest.mark.parametrize("wire", ["plain", "c3", "bm"])
@pytest.mark.parametrize("case", CASES, ids=lambda c: "w{}_b{}_n{}_e{}_c{}_d{}{}".format(*c[:6], "_cap" if len(c) > 6 else ""))
def test_exchange_equals_concatenation_on_the_emulator(emu_bin, wire, case):
    out = subprocess.run([emu_bin, wire] + [str(v) for v in case], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_batches_of_more_than_255_groups_per_cta(emu_bin):
    """2.2 M records per rank on one CTA: 269 groups, i.e. two boundary batches (bm) and the
    non-split search loop (c3)."""
    for wire in ("c3", "bm"):
        out = subprocess.run([emu_bin, wire, "2", "4", "2200000", "2", "1", "33"], capture_output=True, text=True,
                             timeout=900)
        assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


@pytest.mark.parametrize("wire", ["plain", "c3", "bm"])
def test_protocol_survives_skew_between_ranks(emu_bin, wire):
    """EMU_JITTER: random pauses at launches and system-scope stores let ranks drift apart by whole
    kernels over 25 back-to-back ticks; the epoch / done-flag / double-buffer protocol must still
    deliver the concatenation on every rank.  (Deleting the done-flag wait from the kernels makes
    this test fail within two ticks.)"""
    out = subprocess.run([emu_bin, wire, "4", "4", "12000", "25", "3"], capture_output=True, text=True, timeout=600,
                         env=dict(os.environ, EMU_JITTER="1"))
    assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr


def test_flag_protocol_orders_every_cross_rank_access_under_threadsanitizer(tmp_path):
    """The same harness built with -fsanitize=thread (the emulator tells TSan about its fibers; every
    fiber switch synchronises, so only accesses of DIFFERENT ranks can race).  System-scope
    release/acquire are atomics there, everything else — peer payload stores, group counts, bitmaps,
    output reads — plain memory accesses: a report would mean some cross-rank data is not ordered by
    the count / done-flag protocol.  (Without the done-flag wait TSan reports a data race at once.)"""
    exe = str(tmp_path / "emu_gather_tsan.bin")
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-Wno-unknown-pragmas", "-Wno-tsan",
                        "-pthread", "-I", os.path.join(ROOT, "include"), os.path.join(EMU, "emu_gather.cpp"), "-o", exe],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("ThreadSanitizer runtime not available: " + r.stderr[-300:])
    for wire in ("plain", "c3", "bm"):
        out = subprocess.run([exe, wire, "3", "4", "12000", "4", "3"], capture_output=True, text=True, timeout=900,
                             env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0 exitcode=66"))
        assert "ThreadSanitizer" not in out.stderr, out.stderr[-3000:]
        assert out.returncode == 0 and out.stdout.startswith("ok"), out.stdout + out.stderr[-2000:]


def test_bitmap_format_watchdog_gives_up_on_an_absent_peer(emu_bin):
    """Experimental bitmap kernels only: their spin loops are bounded (AMSWEEP_PUSH_TIMEOUT_MS, 200 ms in
    the harness).  A rank that never pushes makes the others report kPeerTimeout in out_counts[world]
    instead of hanging the device."""
    out = subprocess.run([emu_bin, "bm", "3", "4", "20000", "3", "3"], capture_output=True, text=True, timeout=120,
                         env=dict(os.environ, EMU_ABSENT_RANK="1"))
    assert out.returncode == 0 and out.stdout.startswith("ok watchdog"), out.stdout + out.stderr
