"""Multi-GPU plumbing: index-range sharding and the per-tick concatenation of
each GPU's due list.

The record array shards by contiguous global index range, so the global
ascending due list is simply the rank-ordered concatenation of the local
lists.  Status columns never leave their owner GPU; the only exchange per tick
is (1) one count per rank and (2) the variable-length lists.  NCCL has no
allgatherv, so lists are padded to the largest count of the tick.

Backend-agnostic on purpose (torch.distributed: "nccl" on the GPUs, "gloo" in
the CPU tests); nothing here touches record data.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """[first, first+count) of the global index range owned by `rank`: sizes
    differ by at most one, lower ranks take the remainder."""
    q, r = divmod(n_total, world)
    first = rank * q + min(rank, r)
    return first, q + (1 if rank < r else 0)


def allgather_due(idx_local: torch.Tensor, act_local: torch.Tensor, count, shard_base: int,
                  group=None):
    """Concatenate every rank's (local index, action) list in rank order.

    idx_local: int32/int64 tensor of LOCAL indices, at least `count` valid
    entries; act_local: uint8 actions.  Returns (global idx int64[n_total_due],
    action uint8[n_total_due], counts list[int]) on every rank."""
    world = dist.get_world_size(group)
    dev = idx_local.device
    cnt_t = count if isinstance(count, torch.Tensor) else torch.tensor([int(count)], device=dev)
    cnt_t = cnt_t.reshape(1).to(torch.int64)
    counts_t = torch.empty(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts_t, cnt_t, group=group)
    counts = [int(c) for c in counts_t.tolist()]  # host sync: sizes the padded exchange
    maxc = max(counts) if counts else 0
    mine = counts[dist.get_rank(group)]
    if maxc == 0:
        return (torch.empty(0, dtype=torch.int64, device=dev),
                torch.empty(0, dtype=torch.uint8, device=dev), counts)
    pad_idx = torch.zeros(maxc, dtype=torch.int64, device=dev)
    pad_idx[:mine] = idx_local[:mine].to(torch.int64) + shard_base
    pad_act = torch.zeros(maxc, dtype=torch.uint8, device=dev)
    pad_act[:mine] = act_local[:mine]
    all_idx = torch.empty(world * maxc, dtype=torch.int64, device=dev)
    all_act = torch.empty(world * maxc, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(all_idx, pad_idx, group=group)
    dist.all_gather_into_tensor(all_act, pad_act, group=group)
    idx = torch.cat([all_idx[r * maxc: r * maxc + counts[r]] for r in range(world)])
    act = torch.cat([all_act[r * maxc: r * maxc + counts[r]] for r in range(world)])
    return idx, act, counts


class PeerGather:
    """B200-native global due list: per tick every rank ships its sweep's own output (1 bit per
    record + the non-default actions) into every peer over NVLink and rebuilds the global
    (index, action) list locally (csrc/gather.cu: am_gather_exchange), replacing the counts
    all-gather + host sync + padded all-gather above.  torch.distributed is used once, at
    set-up, to swap the CUDA-IPC handles and the shard layout.  `push` is the round-1 list
    format (finished entries on the wire), kept as the measured baseline."""

    def __init__(self, device_index: int, cap_total: int, group=None, idx_bytes: int = 8, shard=None):
        """shard=(first global index, number of records) of this rank: required for `exchange`."""
        import ctypes as C

        from . import _lib as L
        self._L, self._C = L, C
        self._lib = L.load()
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.cap_total = cap_total
        self.device = torch.device("cuda", device_index)
        h = C.c_void_p()
        self.idx_bytes = idx_bytes
        self._h = None

        def agree(ok: bool, what: str):
            """every rank learns whether the step worked everywhere (no rank is left
            waiting in a collective when another one failed)"""
            if self.world > 1:
                flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=self.device)
                dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
                ok = bool(flag.item())
            if not ok:
                self.close()
                raise RuntimeError(f"PeerGather: {what} failed on at least one rank")

        rc = self._lib.am_gather_create(C.byref(h), device_index, self.rank, self.world, cap_total, idx_bytes)
        mine = C.create_string_buffer(L.IPC_HANDLE_BYTES)
        if rc == 0:
            self._h = h
            rc = self._lib.am_gather_export(self._h, mine)
        agree(rc == 0, "am_gather_create/export")
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(mine.raw), group=group)
            rc = self._lib.am_gather_connect(self._h, b"".join(handles))
            agree(rc == 0, "am_gather_connect (CUDA IPC peer mapping)")
        self.has_layout = shard is not None
        if shard is not None:
            import numpy as np
            mine_t = torch.tensor([int(shard[0]), int(shard[1])], dtype=torch.int64, device=self.device)
            all_t = torch.empty(2 * self.world, dtype=torch.int64, device=self.device)
            if self.world > 1:
                dist.all_gather_into_tensor(all_t, mine_t, group=group)
            else:
                all_t.copy_(mine_t)
            lay = all_t.cpu().numpy().astype(np.uint64).reshape(self.world, 2)
            self.bases = np.ascontiguousarray(lay[:, 0])
            self.sizes = np.ascontiguousarray(lay[:, 1])
            rc = self._lib.am_gather_set_layout(self._h, self.bases.ctypes.data, self.sizes.ctypes.data)
            agree(rc == 0, "am_gather_set_layout")

    def _check(self, rc, where):
        if rc != 0:
            raise RuntimeError(f"{where}: {rc}: {self._lib.am_gather_last_error(self._h).decode()}")

    def exchange(self, sweep, d_stats_ptr: int = 0, stream: int = 0):
        """after sweep.tick_shard(): ship this shard's bitmap + exceptions, rebuild the global list"""
        self._check(self._lib.am_gather_exchange(self._h, sweep._h, d_stats_ptr or None, stream or None),
                    "am_gather_exchange")

    def bind(self, sweep, sweep_stream: int = 0, exchange_stream: int = 0):
        """name the shard and the two streams of tick_view()"""
        self._check(self._lib.am_gather_bind(self._h, sweep._h, sweep_stream or None, exchange_stream or None),
                    "am_gather_bind")

    def tick_view(self, unix_sec: int, mode: int = 0):
        """one whole step of the bound shard (tick_shard + exchange + this rank's own part of the global list as
        local slots in the library's pinned host memory): (u32 local idx view, u8 action view, stats of the shard)"""
        import numpy as np
        from . import _lib as L
        C = self._C
        v, st = L.AmTickView(), L.AmTickStats()
        self._check(self._lib.am_gather_tick_view(self._h, unix_sec, mode, C.byref(v), C.byref(st)), "am_gather_tick_view")
        n = int(v.n)
        if n == 0:
            return np.empty(0, np.uint32), np.empty(0, np.uint8), st.as_dict()
        idx = np.ctypeslib.as_array(C.cast(v.idx_local, C.POINTER(C.c_uint32)), shape=(n,))
        act = np.ctypeslib.as_array(C.cast(v.action, C.POINTER(C.c_uint8)), shape=(n,))
        return idx, act, st.as_dict()

    def set_profiling(self, on: bool):
        self._check(self._lib.am_gather_set_profiling(self._h, int(on)), "am_gather_set_profiling")

    def last_profile(self):
        """(push ms incl. the wait for the peers, list rebuild ms, publish ms) of the last exchange"""
        a, b, c = self._C.c_double(), self._C.c_double(), self._C.c_double()
        self._check(self._lib.am_gather_last_profile(self._h, self._C.byref(a), self._C.byref(b), self._C.byref(c)),
                    "am_gather_last_profile")
        return a.value, b.value, c.value

    def push(self, d_idx_ptr: int, d_act_ptr: int, d_count_ptr: int, shard_base: int, stream: int):
        self._check(self._lib.am_gather_push(self._h, d_idx_ptr, d_act_ptr, d_count_ptr, shard_base,
                                             stream or None), "am_gather_push")

    def _wrap(self, ptr: int, n: int, dtype):
        """zero-copy torch view of library-owned device memory"""
        class _Arr:
            pass
        itemsize = torch.empty(0, dtype=dtype).element_size()
        typestr = {torch.int64: "<i8", torch.uint8: "|u1", torch.int32: "<i4"}[dtype]  # u32 viewed as i32
        a = _Arr()
        a.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False),
                                      "version": 3, "strides": (itemsize,)}
        return torch.as_tensor(a, device=self.device)

    def counts(self):
        """device view of out_counts: world per-rank counts, then the total"""
        return self._wrap(self._lib.am_gather_out_counts(self._h), self.world + 1, torch.int32)

    def buffers(self):
        """device views of the CURRENT output buffers, full capacity (the caller slices by the total)"""
        idt = torch.int64 if self.idx_bytes == 8 else torch.int32
        return (self._wrap(self._lib.am_gather_out_idx(self._h), self.cap_total, idt),
                self._wrap(self._lib.am_gather_out_act(self._h), self.cap_total, torch.uint8))

    def result(self):
        """(global idx int64|int32[total], action uint8[total], counts list) — call after
        the exchange / push has retired on its stream (synchronises to read the counts)."""
        counts = self.counts().tolist()
        total = counts[self.world]
        if total & 0xFFFFFFFF == 0xFFFFFFFF:  # the push kernel gave up on a peer
            raise RuntimeError("PeerGather: a peer did not arrive within AMSWEEP_PUSH_TIMEOUT_MS; the handle is "
                               "out of step with its peers and must be recreated")
        idx, act = self.buffers()
        return idx[:total], act[:total], counts[: self.world]

    def close(self):
        if getattr(self, "_h", None):
            self._lib.am_gather_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
