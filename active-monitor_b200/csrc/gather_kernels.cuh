// gather_kernels.cuh — device side of gather.cu: the exchange block header, the kernel
// parameter blocks and the two push kernels (tick exchange: bitmap + exceptions; plain:
// finished lists).  The list rebuild on the receiving side is expand_kernel of
// sweep_kernels.cuh — the same kernel that rebuilds a single GPU's list.  Kept free of
// host API calls so that tests/emu can compile the same source for the CPU (threads as
// fibers, ranks as OS threads) and check the exchange logic without a GPU.
#pragma once
#ifndef AMSWEEP_EMULATE
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "sweep_types.h"  // AM_LAUNCH, TickOut / ExpandParams

namespace {

using amsweep::kMaxWorld;

struct ExchangeHeader {
  unsigned long long count_slot[kMaxWorld];  // {epoch:32 | count:32}, slot r written by rank r
  unsigned long long done_slot[kMaxWorld];   // epoch, slot r written by rank r
  unsigned int cta_done;                     // local CTA ticket
  unsigned int pad[31];
};

struct PushParams {
  unsigned char* peer[kMaxWorld];  // base of every rank's exchange block (own included)
  const uint32_t* idx_local;
  const uint8_t* act_local;
  const uint32_t* count_local;
  uint32_t* out_counts;  // [world+1] on this rank: per-rank counts, then the total
  uint64_t shard_base;
  uint64_t cap_total;
  size_t off_idx[2], off_act[2];  // byte offsets of the two output buffers in a block
  uint32_t epoch;
  int rank, world;
  int idx_bytes;  // 4: u32 global indices on the wire and in the output, 8: u64
};

#ifndef AMSWEEP_EMULATE
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif  // the emulator supplies all three (tests/emu/cuda_emu.h)

constexpr uint32_t kPeerTimeout = 0xFFFFFFFFu;  // out_counts[world] when a peer did not arrive in time

// Bounded wait for a flag word whose upper (count slots) or whole (done slots) value carries the
// epoch.  timeout_ns == 0: wait for ever (the validated formats); otherwise give up after that long,
// so that a peer that died or diverged costs seconds, not a hung GPU.
__device__ __forceinline__ bool wait_flag(const unsigned long long* p, unsigned long long want, bool upper_half,
                                          unsigned long long timeout_ns, unsigned long long* value) {
  const unsigned long long t0 = timeout_ns ? global_timer_ns() : 0ull;
  for (;;) {
    const unsigned long long v = ld_acquire_sys(p);
    if ((upper_half ? (v >> 32) : v) == want) { *value = v; return true; }
    if (timeout_ns && global_timer_ns() - t0 > timeout_ns) { *value = v; return false; }
  }
}

// wait until a peer's done flag has reached `epoch` (epochs are compared modulo 2^32)
__device__ __forceinline__ bool wait_done_at_least(const unsigned long long* p, uint32_t epoch, unsigned long long timeout_ns) {
  const unsigned long long t0 = timeout_ns ? global_timer_ns() : 0ull;
  for (;;) {
    const uint32_t v = (uint32_t)ld_acquire_sys(p);
    if ((int32_t)(v - epoch) >= 0 && v != 0u) return true;
    if (timeout_ns && global_timer_ns() - t0 > timeout_ns) return false;
  }
}

__global__ void __launch_bounds__(256) gather_push_kernel(const PushParams p) {
  __shared__ uint32_t s_count[kMaxWorld];
  const int tid = threadIdx.x;
  ExchangeHeader* mine = reinterpret_cast<ExchangeHeader*>(p.peer[p.rank]);
  const uint32_t my_count = *p.count_local;

  // 1. publish my count to every peer
  if (blockIdx.x == 0 && tid < p.world) {
    ExchangeHeader* peer = reinterpret_cast<ExchangeHeader*>(p.peer[tid]);
    st_release_sys(&peer->count_slot[p.rank], ((unsigned long long)p.epoch << 32) | my_count);
  }
  // 2. wait for every rank's count of this epoch
  if (tid < p.world) {
    unsigned long long v;
    do { v = ld_acquire_sys(&mine->count_slot[tid]); } while ((uint32_t)(v >> 32) != p.epoch);
    s_count[tid] = (uint32_t)v;
  }
  __syncthreads();
  uint64_t offset = 0, total = 0;
  for (int r = 0; r < p.world; ++r) {
    if (r < p.rank) offset += s_count[r];
    total += s_count[r];
  }
  // 3. write my list into every peer's output buffer at `offset`.  Work items are
  //    DESTINATION-aligned quads (4 consecutive output slots): an interior quad is
  //    one 16 B index store + one 4 B action store per peer (NVLink packets of
  //    512 B / 128 B per warp) instead of eight scalar stores; the ragged first and
  //    last quads fall back to scalar stores.  Source reads are local and unaligned
  //    (scalar, L2-resident: the list was just written by expand_kernel).
  //    Four quads are in flight per thread so a small grid saturates the link
  //    while leaving the SMs to the next tick's sweep.
  const int buf = p.epoch & 1;
  const uint64_t room = offset < p.cap_total ? p.cap_total - offset : 0;
  const uint64_t n = my_count < room ? my_count : room;
  const uint64_t q_lo = offset / 4, q_hi = (offset + n + 3) / 4;  // quads [q_lo, q_hi)
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t q0 = q_lo + blockIdx.x * (uint64_t)blockDim.x + tid; q0 < q_hi; q0 += 4 * stride) {
    uint32_t li[4][4], ga[4];  // local indices and packed action bytes of four quads
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t q = q0 + (uint64_t)u * stride;
      ga[u] = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t pos = 4 * q + k;  // output slot
        const bool ok = q < q_hi && pos >= offset && pos < offset + n;
        const uint64_t e = pos - offset;
        li[u][k] = ok ? __ldcs(p.idx_local + e) : 0u;
        ga[u] |= (ok ? (uint32_t)__ldcs(p.act_local + e) : 0u) << (8 * k);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t q = q0 + (uint64_t)u * stride;
      if (q >= q_hi) continue;
      const bool full = 4 * q >= offset && 4 * q + 4 <= offset + n;
      if (full && p.idx_bytes == 4) {
        const uint32_t b32 = (uint32_t)p.shard_base;
        const uint4 v = make_uint4(b32 + li[u][0], b32 + li[u][1], b32 + li[u][2], b32 + li[u][3]);
        for (int r = 0; r < p.world; ++r) {
          reinterpret_cast<uint4*>(p.peer[r] + p.off_idx[buf])[q] = v;
          reinterpret_cast<uint32_t*>(p.peer[r] + p.off_act[buf])[q] = ga[u];
        }
      } else if (full) {
        const ulonglong2 v0 = make_ulonglong2(p.shard_base + li[u][0], p.shard_base + li[u][1]);
        const ulonglong2 v1 = make_ulonglong2(p.shard_base + li[u][2], p.shard_base + li[u][3]);
        for (int r = 0; r < p.world; ++r) {
          ulonglong2* d = reinterpret_cast<ulonglong2*>(p.peer[r] + p.off_idx[buf]) + 2 * q;
          d[0] = v0;
          d[1] = v1;
          reinterpret_cast<uint32_t*>(p.peer[r] + p.off_act[buf])[q] = ga[u];
        }
      } else {  // ragged first / last quad of my segment: scalar stores
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t pos = 4 * q + k;
          if (pos < offset || pos >= offset + n) continue;
          const uint64_t g = p.shard_base + li[u][k];
          for (int r = 0; r < p.world; ++r) {
            if (p.idx_bytes == 4) reinterpret_cast<uint32_t*>(p.peer[r] + p.off_idx[buf])[pos] = (uint32_t)g;
            else reinterpret_cast<uint64_t*>(p.peer[r] + p.off_idx[buf])[pos] = g;
            (p.peer[r] + p.off_act[buf])[pos] = (uint8_t)(ga[u] >> (8 * k));
          }
        }
      }
    }
  }
  // 4. completion: last CTA raises my done flag everywhere, then waits for all peers
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&mine->cta_done, 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence_system();
      for (int r = 0; r < p.world; ++r) {
        ExchangeHeader* peer = reinterpret_cast<ExchangeHeader*>(p.peer[r]);
        st_release_sys(&peer->done_slot[p.rank], (unsigned long long)p.epoch);
      }
      for (int r = 0; r < p.world; ++r) {
        while (ld_acquire_sys(&mine->done_slot[r]) != (unsigned long long)p.epoch) {}
      }
      for (int r = 0; r < p.world; ++r) p.out_counts[r] = s_count[r];
      p.out_counts[p.world] = (uint32_t)(total < p.cap_total ? total : p.cap_total);
      mine->cta_done = 0;
      __threadfence();
    }
  }
}



// ---------------------------------------------------------------------------
// Tick exchange: what crosses NVLink is the sweep's own output — one bit per record for the
// emitted set (1 KB per 8192-record group), the per-group offsets, the per-tile exception
// counts and the exception entries (records whose action is not the bare SUBMIT_HC) — instead
// of finished (index, action) entries: ~1.3 MB per 10 M-record shard per peer at the bench
// density instead of 16.7 MB (round-1 plain format) or 10 MB (round-1 "c3").  Every GPU then
// rebuilds the GLOBAL list from the world's bitmaps with expand_kernel — local HBM writes,
// no wire traffic.  There is no count exchange and therefore no wait before the payload:
// a rank's offset in the global list follows from the group prefixes it receives.
//
// Buffer reuse (two slot sets by epoch parity): a rank finishes push e-1 only after every
// peer raised done(e-1); a peer's push e-1 is stream-ordered after its expand e-2, so when
// this rank starts push e, every peer has finished reading what push e-2 wrote.
// ---------------------------------------------------------------------------
struct PushTickParams {
  unsigned char* peer[kMaxWorld];  // base of every rank's exchange block (own included)
  const uint32_t* bitmap;          // this rank's tick output (amsweep::TickOut)
  const uint32_t* group_prefix;
  const uint32_t* tile_exc;
  const uint32_t* exc_seg;
  size_t off_bitmap, off_prefix, off_tile_exc, off_exc;  // of THIS rank's slot (this parity) in every block
  uint32_t n_groups, n_tiles;
  uint32_t epoch;
  int rank, world;
  unsigned long long timeout_ns;  // 0 = wait for ever; else give up on a peer after this long
  uint32_t* status;               // sticky: set to 1 when a peer did not arrive in time
};

__global__ void __launch_bounds__(256) gather_push_tick_kernel(const PushTickParams p) {
  const int tid = threadIdx.x, lane = tid & 31;
  const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + tid;
  const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
  ExchangeHeader* mine = reinterpret_cast<ExchangeHeader*>(p.peer[p.rank]);

  // A. bitmap, 16 B per store: n_groups * 256 words (the words past the last tile are zero)
  const uint64_t n4 = (uint64_t)p.n_groups * amsweep::kGroupWords / 4;
  const uint4* bm4 = reinterpret_cast<const uint4*>(p.bitmap);
  for (uint64_t i = gtid; i < n4; i += nthreads) {
    const uint4 v = bm4[i];
    for (int r = 0; r < p.world; ++r)
      if (r != p.rank) reinterpret_cast<uint4*>(p.peer[r] + p.off_bitmap)[i] = v;
  }
  // B. group offsets and per-tile exception counts
  for (uint64_t i = gtid; i < (uint64_t)p.n_groups + 1; i += nthreads) {
    const uint32_t v = p.group_prefix[i];
    for (int r = 0; r < p.world; ++r)
      if (r != p.rank) reinterpret_cast<uint32_t*>(p.peer[r] + p.off_prefix)[i] = v;
  }
  for (uint64_t i = gtid; i < (uint64_t)p.n_tiles; i += nthreads) {
    const uint32_t v = p.tile_exc[i];
    for (int r = 0; r < p.world; ++r)
      if (r != p.rank) reinterpret_cast<uint32_t*>(p.peer[r] + p.off_tile_exc)[i] = v;
  }
  // C. exception entries: one warp per tile copies the used prefix of the tile's segment
  const uint64_t nwarps = nthreads >> 5;
  for (uint64_t t = gtid >> 5; t < (uint64_t)p.n_tiles; t += nwarps) {
    const uint32_t nx = p.tile_exc[t];
    for (uint32_t i = (uint32_t)lane; i < nx; i += 32u) {
      const size_t e = (size_t)t * amsweep::kTile + i;
      const uint32_t v = p.exc_seg[e];
      for (int r = 0; r < p.world; ++r)
        if (r != p.rank) reinterpret_cast<uint32_t*>(p.peer[r] + p.off_exc)[e] = v;
    }
  }
  // D. completion: last CTA raises my done flag everywhere, then waits for all peers
  __threadfence_system();
  __syncthreads();
  if (tid == 0) {
    const unsigned int ticket = atomicAdd(&mine->cta_done, 1u);
    if (ticket == gridDim.x - 1) {
      __threadfence_system();
      for (int r = 0; r < p.world; ++r) {
        ExchangeHeader* peer = reinterpret_cast<ExchangeHeader*>(p.peer[r]);
        st_release_sys(&peer->done_slot[p.rank], (unsigned long long)p.epoch);
      }
      // ">= epoch", not "== epoch": there is no count exchange at the start of this kernel, so a fast
      // peer may already have finished its NEXT push (other parity) and raised done(epoch + 1) while
      // this rank is still waiting here
      for (int r = 0; r < p.world; ++r)
        if (!wait_done_at_least(&mine->done_slot[r], p.epoch, p.timeout_ns)) *p.status = 1u;
      mine->cta_done = 0;
      __threadfence();
    }
  }
}

// per-rank counts and the total of an exchanged tick, from the group prefixes (one thread)
struct CountsParams {
  const uint32_t* group_prefix[kMaxWorld];
  uint32_t n_groups[kMaxWorld];
  uint32_t* out_counts;  // [world + 1]
  const uint32_t* status;
  uint64_t cap_total;
  int world;
};
__global__ void gather_counts_kernel(const CountsParams p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (*p.status) {  // a peer timed out: the slots hold an older tick
    for (int r = 0; r < p.world; ++r) p.out_counts[r] = 0;
    p.out_counts[p.world] = kPeerTimeout;
    return;
  }
  uint64_t total = 0;
  for (int r = 0; r < p.world; ++r) {
    const uint32_t c = p.group_prefix[r][p.n_groups[r]];
    p.out_counts[r] = c;
    total += c;
  }
  p.out_counts[p.world] = (uint32_t)(total < p.cap_total ? total : p.cap_total);
}

// This rank's own part of the rebuilt global list — its checks, as LOCAL slots — into (mapped host) memory:
// the multi-GPU counterpart of am_sweep_tick_view.  The part's offset follows from the counts on the device,
// so the host needs no count before the copy and synchronises once.
struct ExtractParams {
  const void* gidx;          // global list, u32 or u64
  const uint8_t* gact;
  const uint32_t* counts;    // [world + 1]
  uint32_t* out_idx;         // [cap] local slots
  uint8_t* out_act;          // [cap]
  unsigned long long* out_n; // entries written (kPeerTimeout << 32 when a peer timed out)
  uint64_t shard_base, cap;
  int rank, world, idx_bytes;
};
__global__ void gather_extract_own_kernel(const ExtractParams p) {
  uint64_t off = 0;
  for (int r = 0; r < p.rank; ++r) off += p.counts[r];
  const bool timed_out = p.counts[p.world] == kPeerTimeout;
  uint64_t mine = timed_out ? 0 : p.counts[p.rank];
  if (mine > p.cap) mine = p.cap;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < mine; k += stride) {
    const uint64_t g = p.idx_bytes == 4 ? (uint64_t)reinterpret_cast<const uint32_t*>(p.gidx)[off + k]
                                        : reinterpret_cast<const uint64_t*>(p.gidx)[off + k];
    // (u32 global indices wrap at 2^32 exactly as the shard base does: the difference is the local slot)
    p.out_idx[k] = p.idx_bytes == 4 ? (uint32_t)g - (uint32_t)p.shard_base : (uint32_t)(g - p.shard_base);
    p.out_act[k] = p.gact[off + k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) *p.out_n = timed_out ? ((unsigned long long)kPeerTimeout << 32) : mine;
}

}  // namespace
