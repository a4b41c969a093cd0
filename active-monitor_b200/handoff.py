"""Hand-off queue between the tick loop and the workflow workers.

ctypes face of `am_handoff_*` (include/amsweep.h): the ticker publishes the list
`Sweep.tick()` returned, up to MaxParallel workers (hcc.go:138, :298) pop chunks
and do what createSubmitWorkflow / processRemedy (hcc.go:502, :759) do.  Host
only; ctypes releases the GIL during the calls, so Python threads exercise the
same locking a cgo caller would.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .sweep import AmError


class Handoff:
    def __init__(self, capacity: int):
        self._lib = L.load()
        h = C.c_void_p()
        rc = self._lib.am_handoff_create(C.byref(h), capacity)
        if rc != L.AM_OK:
            raise AmError(rc, "am_handoff_create")
        self._h = h
        self.capacity = capacity

    def close(self) -> None:
        if self._h:
            self._lib.am_handoff_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def publish(self, unix_sec: int, idx, action, mask: int = L.ACT_SUBMIT_HC | L.ACT_RUN_REMEDY) -> int:
        """Enqueue one tick's entries whose action intersects `mask`; returns how many.
        Raises AmError(AM_E_NOSPACE) (nothing enqueued) when they do not fit."""
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        action = np.ascontiguousarray(action, dtype=np.uint32)
        if idx.shape != action.shape:
            raise ValueError("idx and action must have the same length")
        n_out = L.u64(0)
        rc = self._lib.am_handoff_publish(self._h, unix_sec, idx.size, idx.ctypes.data,
                                          action.ctypes.data, mask, C.byref(n_out))
        if rc != L.AM_OK:
            raise AmError(rc, "am_handoff_publish", f"{n_out.value} items do not fit")
        return int(n_out.value)

    def pop(self, max_items: int) -> np.ndarray:
        """Up to `max_items` work items in FIFO order (structured array, maybe empty)."""
        out = np.empty(max_items, dtype=L.WORK_ITEM_DTYPE)
        n_out = L.u64(0)
        rc = self._lib.am_handoff_pop(self._h, max_items, out.ctypes.data, C.byref(n_out))
        if rc != L.AM_OK:
            raise AmError(rc, "am_handoff_pop")
        return out[: n_out.value]

    def stats(self) -> dict:
        v = [L.u64(0) for _ in range(4)]
        rc = self._lib.am_handoff_stats(self._h, *[C.byref(x) for x in v])
        if rc != L.AM_OK:
            raise AmError(rc, "am_handoff_stats")
        return dict(zip(("pending", "published", "popped", "rejected_batches"), (int(x.value) for x in v)))
