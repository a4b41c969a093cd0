// gather_kernels.cuh — device side of gather.cu: the exchange block header, the kernel
// parameter blocks and the push / expand kernels of the three wire formats.  Kept free of
// host API calls so that tests/emu can compile the same source for the CPU (threads as
// fibers, ranks as OS threads) and check the exchange logic without a GPU.
#pragma once
#ifndef AMSWEEP_EMULATE
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/amsweep.h"

// one spelling for kernel launches (see sweep_kernels.cuh): <<<>>> under nvcc, emu::launch on the emulator
#ifndef AM_LAUNCH
#ifndef AMSWEEP_EMULATE
#define AM_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#else
#define AM_LAUNCH(kernel, grid, block, stream, ...) ((void)(stream), emu::launch(kernel, dim3(grid), dim3(block), __VA_ARGS__))
#endif
#endif

namespace {

constexpr int kMaxWorld = 16;

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
  //    (scalar, L2-resident: the list was just written by compact_kernel).
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
// Compressed wire format ("c3"): 3 bytes per entry over NVLink instead of 5.
// A rank's list holds ascending LOCAL indices, so within a group of kGroupRecords
// = 8192 consecutive records an entry is a 13-bit offset: the sender ships u16
// offsets + u8 actions (destination-aligned 8 B + 4 B stores per quad) and one
// count per group; every receiver scans the counts and expands
//   global index = base[rank] + 8192 * group + offset
// locally into its final index list (the action bytes already sit in their final
// place; the prefix sum over the group counts is folded into the expansion kernel).
// The expansion costs 6 B of local HBM traffic per entry, the exchange saves 2 B of
// NVLink traffic per entry per peer — NVLink is the scarce resource.
// ---------------------------------------------------------------------------
constexpr uint32_t kGroupRecords = 8192;

struct PushC3Params {
  unsigned char* peer[kMaxWorld];
  const uint32_t* idx_local;
  const uint8_t* act_local;
  const uint32_t* count_local;
  uint32_t* out_counts;
  uint64_t cap_total;
  size_t off_act[2], off_gc[2], off_o16[2];
  uint32_t epoch;
  uint32_t ngroups_mine;  // groups of this rank's shard
  uint32_t ngroups_max;   // row length of the group-count table
  int rank, world;
};

// first position in the ascending list whose index is >= key
__device__ __forceinline__ uint32_t lower_bound_idx(const uint32_t* a, uint32_t n, uint32_t key) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

__global__ void __launch_bounds__(256) gather_push_c3_kernel(const PushC3Params p) {
  __shared__ uint32_t s_count[kMaxWorld];
  const int tid = threadIdx.x;
  ExchangeHeader* mine = reinterpret_cast<ExchangeHeader*>(p.peer[p.rank]);
  const uint32_t my_count = *p.count_local;
  const int buf = p.epoch & 1;

  if (blockIdx.x == 0 && tid < p.world) {  // 1. publish my count
    ExchangeHeader* peer = reinterpret_cast<ExchangeHeader*>(p.peer[tid]);
    st_release_sys(&peer->count_slot[p.rank], ((unsigned long long)p.epoch << 32) | my_count);
  }
  if (tid < p.world) {  // 2. everyone's counts
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
  const uint64_t room = offset < p.cap_total ? p.cap_total - offset : 0;
  const uint32_t n = (uint32_t)(my_count < room ? my_count : room);

  // 2b. my per-group counts -> row `rank` of every peer's table.  The last S CTAs of the
  //     grid do nothing else: each of their threads finds ONE group boundary by binary
  //     search on the ascending list (255 groups per CTA: 256 boundaries), counts are
  //     differences of neighbours.  The other CTAs go straight to the payload, so the
  //     ~22 dependent L2 reads of a search overlap the NVLink stores.
  __shared__ uint32_t s_bound[256];
  const uint32_t n_search = (p.ngroups_mine + 254u) / 255u;
  const bool split = gridDim.x > n_search;            // enough CTAs to dedicate some
  const uint32_t n_work = split ? gridDim.x - n_search : gridDim.x;
  const bool searcher = split ? blockIdx.x >= n_work : true;
  if (searcher) {
    const uint32_t first_sb = split ? blockIdx.x - n_work : blockIdx.x;
    for (uint32_t sb = first_sb; sb < n_search; sb += (split ? n_search : gridDim.x)) {
      const uint32_t g0 = sb * 255u;
      const uint32_t gb = g0 + (uint32_t)tid;  // boundary index
      // the boundary after the last group is n by definition (and gb * 8192 could wrap there)
      s_bound[tid] = gb < p.ngroups_mine ? lower_bound_idx(p.idx_local, n, gb * kGroupRecords) : n;
      __syncthreads();
      if (tid < 255 && g0 + (uint32_t)tid < p.ngroups_mine) {
        const uint32_t cnt = s_bound[tid + 1] - s_bound[tid];
        for (int r = 0; r < p.world; ++r)
          reinterpret_cast<uint32_t*>(p.peer[r] + p.off_gc[buf])[(size_t)p.rank * p.ngroups_max + g0 + tid] = cnt;
      }
      __syncthreads();
    }
  }
  const bool worker = split ? blockIdx.x < n_work : true;
  const uint64_t stride = (uint64_t)n_work * blockDim.x;

  // 3. offsets + actions, destination-aligned quads: one 8 B and one 4 B store per peer
  const uint64_t q_lo = offset / 4, q_hi = worker ? (offset + n + 3) / 4 : 0;
  for (uint64_t q0 = q_lo + blockIdx.x * (uint64_t)blockDim.x + tid; q0 < q_hi; q0 += 4 * stride) {
    uint32_t lo16[4][2], ga[4];  // two u16 offsets per word, four action bytes per word
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t q = q0 + (uint64_t)u * stride;
      ga[u] = 0; lo16[u][0] = 0; lo16[u][1] = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t pos = 4 * q + k;
        const bool ok = q < q_hi && pos >= offset && pos < offset + n;
        const uint64_t e = pos - offset;
        const uint32_t o = ok ? (__ldcs(p.idx_local + e) & (kGroupRecords - 1u)) : 0u;
        lo16[u][k >> 1] |= o << (16 * (k & 1));
        ga[u] |= (ok ? (uint32_t)__ldcs(p.act_local + e) : 0u) << (8 * k);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t q = q0 + (uint64_t)u * stride;
      if (q >= q_hi) continue;
      const bool full = 4 * q >= offset && 4 * q + 4 <= offset + n;
      if (full) {
        const uint2 v = make_uint2(lo16[u][0], lo16[u][1]);
        for (int r = 0; r < p.world; ++r) {
          reinterpret_cast<uint2*>(p.peer[r] + p.off_o16[buf])[q] = v;
          reinterpret_cast<uint32_t*>(p.peer[r] + p.off_act[buf])[q] = ga[u];
        }
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t pos = 4 * q + k;
          if (pos < offset || pos >= offset + n) continue;
          const uint16_t o = (uint16_t)(lo16[u][k >> 1] >> (16 * (k & 1)));
          for (int r = 0; r < p.world; ++r) {
            reinterpret_cast<uint16_t*>(p.peer[r] + p.off_o16[buf])[pos] = o;
            (p.peer[r] + p.off_act[buf])[pos] = (uint8_t)(ga[u] >> (8 * k));
          }
        }
      }
    }
  }
  // 4. completion, as in the plain format
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

// Receiver: expand one (rank, group) per CTA into the final index list.  The CTA's
// start position is the rank's offset plus the sum of that rank's earlier group counts
// (a few KB of L2-resident reads, as in compact_kernel); four entries are in flight per
// thread.
struct DecodeParams {
  const uint16_t* o16;       // [cap_total] in my exchange block
  const uint32_t* gc;        // [world][ngroups_max] in my exchange block
  const uint32_t* counts;    // out_counts: per-rank counts
  void* final_idx;           // [cap_total] u32 or u64
  uint64_t bases[kMaxWorld];
  uint32_t ngroups[kMaxWorld];
  uint64_t cap_total;
  uint32_t ngroups_max;
  int world, idx_bytes;
};
__global__ void __launch_bounds__(256) gather_decode_kernel(const DecodeParams p) {
  __shared__ uint32_t s_part[8];
  const int r = blockIdx.y, tid = threadIdx.x;
  const uint32_t g = blockIdx.x;
  if (g >= p.ngroups[r]) return;  // uniform per CTA
  const uint32_t* row = p.gc + (size_t)r * p.ngroups_max;
  uint32_t part = 0;
  for (uint32_t j = tid; j < g; j += blockDim.x) part += row[j];
  part = __reduce_add_sync(0xFFFFFFFFu, part);
  if ((tid & 31) == 0) s_part[tid >> 5] = part;
  __syncthreads();
  uint64_t start = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) start += s_part[k];
  for (int q = 0; q < r; ++q) start += p.counts[q];
  const uint32_t cnt = row[g];
  const uint64_t gbase = p.bases[r] + (uint64_t)g * kGroupRecords;
  for (uint32_t e0 = tid; e0 < cnt; e0 += 4u * blockDim.x) {
    uint32_t o[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t e = e0 + (uint32_t)u * blockDim.x;
      o[u] = (e < cnt && start + e < p.cap_total) ? (uint32_t)__ldcs(p.o16 + start + e) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t e = e0 + (uint32_t)u * blockDim.x;
      if (e >= cnt || start + e >= p.cap_total) continue;
      const uint64_t v = gbase + o[u];
      if (p.idx_bytes == 4) reinterpret_cast<uint32_t*>(p.final_idx)[start + e] = (uint32_t)v;
      else reinterpret_cast<uint64_t*>(p.final_idx)[start + e] = v;
    }
  }
}


// ---------------------------------------------------------------------------
// Bitmap wire format ("bm") — EXPERIMENTAL: written after the round-1 GPU budget was
// spent, compiled but not yet executed on hardware; nothing selects it by default.
//
// The NVLink volume, not the sweep, bounds the step from 4 GPUs up (profiles/
// r01_scaling.md), and the payload is far more compressible than c3's 3 B/entry:
//   * the emitted SET of a shard is one bit per record (1 KB per 8192-record group):
//     at 33 % density 3 bits per entry instead of 16;
//   * almost every action byte is the bare AM_ACT_SUBMIT_HC: every receiver pre-fills
//     its action list with that default (stream-ordered before its own push, hence
//     before any peer can write this epoch) and senders skip destination quads that
//     hold nothing else.  A workload of non-default actions degrades to 1 B/entry.
// Receivers rebuild the global index list from the eight bitmaps with a popcount
// prefix per group (gather_expand_bitmap_kernel), starting each group at the sum of
// the sender's earlier group counts, as in c3.
// ---------------------------------------------------------------------------
constexpr uint32_t kGroupWords = kGroupRecords / 32;  // 256 bitmap words per group
constexpr uint32_t kDefaultAction4 = 0x01010101u * AM_ACT_SUBMIT_HC;

struct PushBmParams {
  unsigned char* peer[kMaxWorld];
  const uint32_t* idx_local;
  const uint8_t* act_local;
  const uint32_t* count_local;
  uint32_t* out_counts;
  uint64_t cap_total;
  size_t off_act[2], off_gc[2], off_bm[2];
  uint64_t bm_word0;      // first bitmap word of this rank's row (sum of earlier ranks' groups * 256)
  unsigned long long timeout_ns;  // 0 = wait for ever; else give up on a peer after this long
  uint32_t epoch;
  uint32_t ngroups_mine;
  uint32_t ngroups_max;
  int rank, world;
};

__global__ void __launch_bounds__(256) gather_push_bm_kernel(const PushBmParams p) {
  __shared__ uint32_t s_count[kMaxWorld];
  __shared__ uint32_t s_bound[256];
  __shared__ uint32_t s_bm[kGroupWords];
  const int tid = threadIdx.x;
  ExchangeHeader* mine = reinterpret_cast<ExchangeHeader*>(p.peer[p.rank]);
  const uint32_t my_count = *p.count_local;
  const int buf = p.epoch & 1;

  if (blockIdx.x == 0 && tid < p.world) {  // 1. publish my count
    ExchangeHeader* peer = reinterpret_cast<ExchangeHeader*>(p.peer[tid]);
    st_release_sys(&peer->count_slot[p.rank], ((unsigned long long)p.epoch << 32) | my_count);
  }
  __shared__ uint32_t s_late;  // a peer's count did not arrive in time: nothing is sent
  if (tid == 0) s_late = 0;
  __syncthreads();
  if (tid < p.world) {  // 2. everyone's counts
    unsigned long long v;
    if (!wait_flag(&mine->count_slot[tid], p.epoch, true, p.timeout_ns, &v)) { v = 0; s_late = 1; }
    s_count[tid] = (uint32_t)v;
  }
  __syncthreads();
  const bool late = s_late != 0;
  uint64_t offset = 0, total = 0;
  for (int r = 0; r < p.world; ++r) {
    if (r < p.rank) offset += s_count[r];
    total += s_count[r];
  }
  const uint64_t room = offset < p.cap_total ? p.cap_total - offset : 0;
  const uint32_t n = late ? 0u : (uint32_t)(my_count < room ? my_count : room);

  // 3a. bitmaps + group counts.  Every CTA owns a contiguous run of this shard's groups,
  //     taken in batches of up to 255: 256 threads find the batch's boundaries in the
  //     ascending list (one binary search each, in parallel), then the CTA builds one
  //     group's 256-word bitmap at a time in shared memory and stores it to every rank.
  const uint32_t per_cta = (p.ngroups_mine + gridDim.x - 1) / gridDim.x;
  const uint32_t g_first = blockIdx.x * per_cta;
  const uint32_t g_last = g_first + per_cta < p.ngroups_mine ? g_first + per_cta : p.ngroups_mine;
  for (uint32_t b0 = g_first; b0 < g_last; b0 += 255u) {
    const uint32_t nb = g_last - b0 < 255u ? g_last - b0 : 255u;  // groups in this batch
    if ((uint32_t)tid <= nb) {
      const uint32_t gb = b0 + (uint32_t)tid;
      s_bound[tid] = gb < p.ngroups_mine ? lower_bound_idx(p.idx_local, n, gb * kGroupRecords) : n;
    }
    __syncthreads();
    if ((uint32_t)tid < nb) {
      const uint32_t cnt = s_bound[tid + 1] - s_bound[tid];
      for (int r = 0; r < p.world; ++r)
        reinterpret_cast<uint32_t*>(p.peer[r] + p.off_gc[buf])[(size_t)p.rank * p.ngroups_max + b0 + tid] = cnt;
    }
    for (uint32_t k = 0; k < nb; ++k) {
      s_bm[tid] = 0u;  // blockDim.x == kGroupWords
      __syncthreads();
      const uint32_t lo = s_bound[k], hi = s_bound[k + 1];
      for (uint32_t e = lo + (uint32_t)tid; e < hi; e += blockDim.x) {
        const uint32_t o = __ldcs(p.idx_local + e) & (kGroupRecords - 1u);
        atomicOr(&s_bm[o >> 5], 1u << (o & 31u));
      }
      __syncthreads();
      const uint32_t w = s_bm[tid];
      const uint64_t word = p.bm_word0 + (uint64_t)(b0 + k) * kGroupWords + (uint64_t)tid;
      for (int r = 0; r < p.world; ++r)
        reinterpret_cast<uint32_t*>(p.peer[r] + p.off_bm[buf])[word] = w;
      __syncthreads();  // s_bm is re-zeroed by the next group
    }
  }

  // 3b. actions, destination-aligned quads as in the other formats, except that a quad (or a
  //     ragged byte) holding only the default action is not sent: the receiver pre-filled it.
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t q_lo = offset / 4, q_hi = (offset + n + 3) / 4;
  for (uint64_t q0 = q_lo + blockIdx.x * (uint64_t)blockDim.x + tid; q0 < q_hi; q0 += 4 * stride) {
    uint32_t ga[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t q = q0 + (uint64_t)u * stride;
      ga[u] = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t pos = 4 * q + k;
        const bool ok = q < q_hi && pos >= offset && pos < offset + n;
        ga[u] |= (ok ? (uint32_t)__ldcs(p.act_local + (pos - offset)) : 0u) << (8 * k);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint64_t q = q0 + (uint64_t)u * stride;
      if (q >= q_hi) continue;
      const bool full = 4 * q >= offset && 4 * q + 4 <= offset + n;
      if (full) {
        if (ga[u] == kDefaultAction4) continue;
        for (int r = 0; r < p.world; ++r)
          reinterpret_cast<uint32_t*>(p.peer[r] + p.off_act[buf])[q] = ga[u];
      } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t pos = 4 * q + k;
          if (pos < offset || pos >= offset + n) continue;
          const uint8_t a = (uint8_t)(ga[u] >> (8 * k));
          if (a == (uint8_t)AM_ACT_SUBMIT_HC) continue;
          for (int r = 0; r < p.world; ++r) (p.peer[r] + p.off_act[buf])[pos] = a;
        }
      }
    }
  }
  // 4. completion, as in the plain format
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
      bool all_done = !late;
      for (int r = 0; r < p.world; ++r) {
        unsigned long long v;
        if (!wait_flag(&mine->done_slot[r], (unsigned long long)p.epoch, false, p.timeout_ns, &v)) all_done = false;
      }
      for (int r = 0; r < p.world; ++r) p.out_counts[r] = s_count[r];
      p.out_counts[p.world] = all_done ? (uint32_t)(total < p.cap_total ? total : p.cap_total) : kPeerTimeout;
      mine->cta_done = 0;
      __threadfence();
    }
  }
}

// Receiver: one CTA per (rank, group) turns the group's 256 bitmap words back into
// ascending global indices.  Start position as in gather_decode_kernel; in-group rank of
// a word = exclusive prefix of the popcounts (warp shuffles + 8 warp totals); the indices
// are staged in shared memory so that the stores to the final list are coalesced.
struct ExpandBmParams {
  const uint32_t* bm;        // bitmap area of my exchange block (all ranks' rows)
  const uint32_t* gc;        // [world][ngroups_max]
  const uint32_t* counts;    // out_counts
  void* final_idx;
  uint64_t bases[kMaxWorld];
  uint64_t bm_word0[kMaxWorld];
  uint32_t ngroups[kMaxWorld];
  uint64_t cap_total;
  uint32_t ngroups_max;
  int world, idx_bytes;
};
__global__ void __launch_bounds__(256) gather_expand_bitmap_kernel(const ExpandBmParams p) {
  __shared__ uint32_t s_part[8];
  __shared__ uint32_t s_warp[8];
  __shared__ uint32_t s_off[kGroupRecords];  // in-group offsets of the set bits, ascending
  const int r = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t g = blockIdx.x;
  if (g >= p.ngroups[r]) return;  // uniform per CTA
  if (p.counts[p.world] == kPeerTimeout) return;  // the push gave up on a peer: nothing to expand
  const uint32_t* row = p.gc + (size_t)r * p.ngroups_max;
  uint32_t part = 0;
  for (uint32_t j = tid; j < g; j += blockDim.x) part += row[j];
  part = __reduce_add_sync(0xFFFFFFFFu, part);
  uint32_t w = __ldcs(p.bm + p.bm_word0[r] + (uint64_t)g * kGroupWords + (uint64_t)tid);
  const uint32_t c = (uint32_t)__popc(w);
  uint32_t inc = c;  // inclusive scan of the popcounts within the warp
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, inc, d);
    if (lane >= d) inc += v;
  }
  if (lane == 0) s_part[warp] = part;
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  uint64_t start = 0;
  uint32_t before = 0, cnt = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    start += s_part[k];
    before += k < warp ? s_warp[k] : 0u;
    cnt += s_warp[k];
  }
  for (int q = 0; q < r; ++q) start += p.counts[q];
  uint32_t pos = before + inc - c;  // exclusive prefix of this thread's word
  while (w) {
    const uint32_t b = (uint32_t)__ffs((int)w) - 1u;
    w &= w - 1u;
    s_off[pos++] = (uint32_t)tid * 32u + b;
  }
  __syncthreads();
  const uint64_t gbase = p.bases[r] + (uint64_t)g * kGroupRecords;
  for (uint32_t e = tid; e < cnt; e += blockDim.x) {
    if (start + e >= p.cap_total) break;
    const uint64_t v = gbase + s_off[e];
    if (p.idx_bytes == 4) reinterpret_cast<uint32_t*>(p.final_idx)[start + e] = (uint32_t)v;
    else reinterpret_cast<uint64_t*>(p.final_idx)[start + e] = v;
  }
}

}  // namespace
