"""Child process of tests/test_c_abi_on_emulator.py: `world` ranks as Python threads drive the
REAL am_gather_* C-ABI of the emulated library (csrc/gather.cu host code + gather_kernels.cuh on
cuda_emu.h; CUDA IPC handles carry plain pointers, cuda_rt_emu.h).  ctypes releases the GIL during
a call, so the ranks' push kernels really run concurrently and meet through their flags.

usage: run_gather_ranks.py <wire: plain|c3|bm> <world> <idx_bytes> <records per rank> <ticks>
Prints "ok ..." and exits 0 when every rank's output of every tick equals the concatenation."""
import ctypes as C
import importlib
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
abi = importlib.import_module("active-monitor_b200._lib")
assert "emu" in os.path.basename(abi.LIB_PATH), "run with AMSWEEP_LIB pointing at libamsweep_emu.so"
lib = abi.load()

wire, world, idx_bytes, n_rec, ticks = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
sizes = [n_rec + (37 if r == world - 1 else 0) for r in range(world)]
bases = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.uint64)
cap_total = int(sum(sizes))


def local_list(rank, tick):
    rng = np.random.default_rng(1000 * tick + rank)
    emitted = rng.random(sizes[rank]) < (0.33 if (tick + rank) % 3 else 0.9)
    if tick % 4 == 3 and rank == 1:
        emitted[:] = False
    idx = np.flatnonzero(emitted).astype(np.uint32)
    act = np.where(rng.random(len(idx)) < 0.97, 0x01, rng.choice([0x08, 0x23, 0x80], len(idx))).astype(np.uint8)
    return idx, act


handles = [None] * world
errors = []
bar = threading.Barrier(world)


def rank_main(rank):
    try:
        h = C.c_void_p()
        assert lib.am_gather_create(C.byref(h), rank, rank, world, cap_total, idx_bytes) == 0
        mine = C.create_string_buffer(abi.IPC_HANDLE_BYTES)
        assert lib.am_gather_export(h, mine) == 0
        handles[rank] = bytes(mine.raw)
        bar.wait()
        assert lib.am_gather_connect(h, b"".join(handles)) == 0
        if wire in ("c3", "bm"):
            b = np.ascontiguousarray(bases, dtype=np.uint64)
            s = np.array(sizes, dtype=np.uint64)
            assert lib.am_gather_set_layout(h, b.ctypes.data, s.ctypes.data) == 0
            if wire == "bm":
                assert lib.am_gather_set_wire(h, abi.WIRE_BITMAP) == 0
        for t in range(1, ticks + 1):  # no barrier between ticks, as on a stream
            idx, act = local_list(rank, t)
            cnt = np.array([len(idx)], dtype=np.uint32)
            idx_buf = np.concatenate([idx, np.zeros(4, np.uint32)])  # device buffers are larger than the list
            act_buf = np.concatenate([act, np.zeros(4, np.uint8)])
            rc = lib.am_gather_push(h, idx_buf.ctypes.data, act_buf.ctypes.data, cnt.ctypes.data, int(bases[rank]), None)
            assert rc == 0, (rc, lib.am_gather_last_error(h))
            counts = np.ctypeslib.as_array(C.cast(lib.am_gather_out_counts(h), C.POINTER(C.c_uint32)), (world + 1,)).copy()
            total = int(counts[world])
            it = C.c_uint32 if idx_bytes == 4 else C.c_uint64
            got_idx = np.ctypeslib.as_array(C.cast(lib.am_gather_out_idx(h), C.POINTER(it)), (max(total, 1),))[:total].astype(np.uint64)
            got_act = np.ctypeslib.as_array(C.cast(lib.am_gather_out_act(h), C.POINTER(C.c_uint8)), (max(total, 1),))[:total].copy()
            want_i, want_a = [], []
            for r in range(world):
                li, la = local_list(r, t)
                assert int(counts[r]) == len(li), (rank, t, r, int(counts[r]), len(li))
                want_i.append(li.astype(np.uint64) + bases[r])
                want_a.append(la)
            np.testing.assert_array_equal(got_idx, np.concatenate(want_i), err_msg=f"rank {rank} tick {t} idx")
            np.testing.assert_array_equal(got_act, np.concatenate(want_a), err_msg=f"rank {rank} tick {t} act")
        # argument errors are reported before anything a peer could observe happens
        if wire in ("c3", "bm"):
            assert lib.am_gather_push(h, idx_buf.ctypes.data, act_buf.ctypes.data, cnt.ctypes.data,
                                      int(bases[rank]) + 1, None) == abi.AM_E_INVAL
        bar.wait()
        lib.am_gather_destroy(h)
    except BaseException as e:  # noqa: BLE001
        errors.append((rank, repr(e)))
        try:
            bar.abort()
        except Exception:
            pass


th = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
for t in th:
    t.start()
for t in th:
    t.join()
if errors:
    print("FAILED", errors)
    sys.exit(1)
print(f"ok {wire} world={world} idx_bytes={idx_bytes} records={n_rec} ticks={ticks}")
