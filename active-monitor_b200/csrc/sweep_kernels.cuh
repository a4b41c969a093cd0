// sweep_kernels.cuh — the per-tick schedule-evaluation sweep for sm_100a.
//
// One launch evaluates every HealthCheck record of a shard at one wall-clock
// second T.  It replaces, for N records at once, the
// per-CR decisions of the reference controller (hcc.go = internal/controllers/
// healthcheck_controller.go):
//     hcc.go:227        Workflow.Resource == nil      -> record skipped
//     hcc.go:238-250    pause rule                    -> AM_ACT_STOPPED
//     hcc.go:251-263    cron arm (robfig bitmasks)    -> matches(T), A.7
//     hcc.go:264, :751  interval / timer              -> T - finishedAt >= ras
//     hcc.go:635-661    Succeeded transition          -> apply_result()
//     hcc.go:662-722    Failed transition + remedy gate
//     hcc.go:821-851    remedy result counters
//
// Shape of the kernel (pure integer work, HBM-bandwidth bound, no tensor cores):
//   * records live as SoA columns in HBM; a CTA of 256 threads owns one tile of
//     1024 consecutive records, a warp owns 128 of them;
//   * every warp-level load instruction is fully coalesced: lane L reads the
//     16 B (two u64 records) or 8 B (two i32 records) at column + (w + 2L),
//     twice per tile ("halves"), so 16 independent loads are in flight per
//     lane before the first use (a __syncwarp() after them keeps ptxas from
//     delaying half of them).  Staging the tile in shared memory instead —
//     per-lane cp.async, TMA bulk copies one-shot, and a persistent two-stage
//     TMA pipeline — was built, verified and measured 6-15 % slower for this
//     streaming, near-issue-bound kernel;
//   * the tick's broken-down time is computed once per tick as one-hot words
//     and reaches every CTA through the kernel parameters (constant bank); a
//     5-field schedule fires iff minute&M && hour&H && month&Mo && dayMatches
//     (five ANDs, no loop);
//   * remedy/counter columns are touched only by lanes whose record has a
//     posted result (or is due, in closed-loop mode): 56 B/record otherwise;
//   * emitted (index, action) pairs are compacted IN ORDER without any
//     dependency between CTAs: warp ballots + popc give the in-warp rank, a CTA
//     scan the in-tile rank, and the tile writes its entries to its own
//     segment (offset = first record of the tile) plus one count; a second,
//     tiny kernel (compact_kernel) turns segments into the contiguous ascending
//     list.  A single-pass decoupled look-back was measured first and rejected:
//     in-order completion left SM slots idle (profiles/r01a_lookback_sweep_ncu.csv);
//   * per-tick statistics are reduced lane -> warp (redux) -> CTA (shared
//     atomics) -> global (fire-and-forget RED); the kernel boundary before
//     compact_kernel is the only global synchronisation, so the sweep kernel
//     has one __syncthreads, no fences and no completion tickets.
#pragma once
#ifndef AMSWEEP_EMULATE  // tests/emu compiles this file for the CPU (cuda_emu.h supplies the model)
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "../../include/amsweep.h"
#include "civil.h"

// Kernel launches are spelled through one macro so that tests/emu can compile the host
// runtime (sweep.cu) for the CPU emulator as well; under nvcc it is the plain <<<>>> launch.
#ifndef AM_LAUNCH
#ifndef AMSWEEP_EMULATE
#define AM_LAUNCH(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#else
#define AM_LAUNCH(kernel, grid, block, stream, ...) ((void)(stream), emu::launch(kernel, dim3(grid), dim3(block), __VA_ARGS__))
#endif
#endif
#define AM_SWEEP_KERNEL(closed, masks) sweep_tick_kernel<closed, masks>  // one macro argument

namespace amsweep {

#ifndef AM_BLOCK
#define AM_BLOCK 256
#endif
constexpr int kBlock = AM_BLOCK;
constexpr int kWarps = kBlock / 32;
constexpr int kRecPerWarp = 128;             // 2 halves x 32 lanes x 2 records
constexpr int kTile = kWarps * kRecPerWarp;  // 1024 records per CTA
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kNumAcc = 16;  // == number of u64 fields of am_tick_stats_t

struct DevCols {
  uint64_t *minute, *hour, *dom, *month, *dow;
  int32_t* ras;
  uint32_t* flags;
  int64_t* finished_at;
  int32_t *runs_limit, *reset_interval;
  int32_t *success, *failed, *remedy_success, *remedy_failed, *remedy_total;
  int64_t* remedy_finished_at;
};

struct SweepParams {
  DevCols c;
  uint64_t n_records;
  uint64_t shard_base;
  uint64_t seed;
  int64_t T;
  TickWords words;  // T's UTC fields as one-hot words, computed once per tick by the launcher
  uint32_t n_tiles;
  uint32_t mode;
  uint32_t* seg_idx;         // [n_tiles * kTile] per-tile segments of local indices
  uint8_t* seg_act;          // [n_tiles * kTile] ... and action bytes
  uint32_t* tile_count;      // [n_tiles] entries emitted by each tile
  uint32_t* group_count;     // [n_groups] sum of tile_count over kGroupTiles tiles (zero on entry)
  unsigned long long* acc;   // [kNumAcc] statistics accumulators (zero on entry)
};

#ifndef AM_GROUP_TILES
#define AM_GROUP_TILES 8  // build-time knob for experiments (<= 32: one warp scans a group's tile counts)
#endif
constexpr int kGroupTiles = AM_GROUP_TILES;  // tiles per compaction group
static_assert(kGroupTiles >= 1 && kGroupTiles <= 32 && (kGroupTiles & (kGroupTiles - 1)) == 0,
              "the in-group scan is a power-of-two shuffle ladder inside one warp");

struct CompactParams {
  const uint32_t* seg_idx;
  const uint8_t* seg_act;
  const uint32_t* tile_count;
  const uint32_t* group_count;  // this tick's group sums
  uint32_t* group_count_next;   // the other parity: zeroed here for the next tick
  unsigned long long* acc;      // action-bit counts and index checksums are added here
  uint32_t* out_idx;            // [cap] ascending local indices
  uint8_t* out_act;             // [cap]
  uint64_t shard_base;  // for the global-index checksums
  uint32_t n_tiles, n_groups, cap;
};

// ---- streaming loads / stores: every byte is touched once per tick --------
template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }
template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v) { __stcs(p, v); }

// Segment entries are written once by the sweep and read once, a few tens of
// microseconds later, by compact_kernel: keep them L2-resident (evict_last)
// while 560 MB of evict_first column data streams past them.
#ifndef AMSWEEP_EMULATE
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
  return pol;
}
__device__ __forceinline__ void st_keep_u32(uint32_t* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ void st_keep_u8(uint8_t* p, uint32_t v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.u8 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
#else  // CPU emulation: plain stores, the cache policy has no meaning
inline uint64_t l2_evict_last_policy() { return 0; }
inline void st_keep_u32(uint32_t* p, uint32_t v, uint64_t) { *p = v; }
inline void st_keep_u8(uint8_t* p, uint32_t v, uint64_t) { *p = (uint8_t)v; }
#endif

__device__ __forceinline__ uint64_t sm64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t outcome_key(uint64_t seed, uint64_t gidx, uint64_t t) {
  return sm64(sm64(seed ^ sm64(gidx)) + t);
}

// bit i of a 4-bit value -> byte i of a word (0/1 each)
__device__ __forceinline__ uint32_t spread4(uint32_t v) { return ((v & 0xFu) * 0x00204081u) & 0x01010101u; }

// Mutable per-record state carried through the result state machine.
struct RecState {
  uint32_t flags;
  int64_t fa;       // finishedAt
  int32_t s, f;     // SuccessCount, FailedCount
  int32_t rs, rf, rt;  // RemedySuccessCount, RemedyFailedCount, RemedyTotalRuns
  int64_t rfa;      // RemedyFinishedAt, 0 == nil
  int32_t limit, reset;
};

// watchRemedyWorkflow result, hcc.go:821-851
__device__ __forceinline__ void remedy_result(RecState& r, int64_t T, bool ok, uint32_t& res) {
  if (ok) { r.rs = (int32_t)((uint32_t)r.rs + 1u); res += 1u << 16; }
  else    { r.rf = (int32_t)((uint32_t)r.rf + 1u); res += 1u << 24; }
  r.rt = (int32_t)((uint32_t)r.rs + (uint32_t)r.rf);
  r.rfa = T;
}

// Step 1 of B.3: apply a posted workflow result.  Returns action bits; `res`
// counts {ok, fail, remedy_ok, remedy_fail} in its four bytes (a record can
// take two results in one closed-loop tick) for the metrics counters.
__device__ __forceinline__ uint32_t apply_result(RecState& r, int64_t T, uint32_t& res) {
  uint32_t act = 0;
  const uint32_t f = r.flags;
  if (f & AM_F_PENDING_OK) {  // hcc.go:635-661
    r.s = (int32_t)((uint32_t)r.s + 1u);
    r.fa = T;
    res += 1u;
    if ((f & AM_F_HAS_REMEDY) && r.rt >= 1) {  // :649
      r.rt = r.rs = r.rf = 0;
      r.rfa = 0;
      act |= AM_ACT_RESET_ON_PASS;
    }
  } else if (f & AM_F_PENDING_FAIL) {  // hcc.go:662-722
    r.f = (int32_t)((uint32_t)r.f + 1u);
    r.fa = T;
    res += 1u << 8;
    if (f & AM_F_HAS_REMEDY) {  // :677
      bool run = false;
      if (r.limit != 0 && r.reset != 0) {  // :679
        if (r.limit > r.rt) {               // :681
          run = true;
        } else if (r.rfa == 0) {
          act |= AM_ACT_ANOMALY;  // nil RemedyFinishedAt at :690 (N3)
        } else {
          // int(now.Sub(RemedyFinishedAt).Seconds()): Sub saturates at
          // +-9223372036 s, far outside the i32 resetInterval domain
          int64_t d = (int64_t)((uint64_t)T - (uint64_t)r.rfa);
          d = d > 9223372036ll ? 9223372036ll : (d < -9223372036ll ? -9223372036ll : d);
          if ((int64_t)r.reset >= d) {  // :692
            act |= AM_ACT_REMEDY_SKIP;
          } else {  // :695-704
            r.rt = r.rs = r.rf = 0;
            r.rfa = 0;
            act |= AM_ACT_RESET_ON_INTERVAL;
            run = true;
          }
        }
      } else {  // :712-719
        run = true;
      }
      if (run) {
        act |= AM_ACT_RUN_REMEDY;
        if (f & AM_F_REMEDY_PENDING) remedy_result(r, T, (f & AM_F_REMEDY_OUTCOME_OK) != 0, res);
      }
    }
  } else if (f & AM_F_REMEDY_PENDING) {  // remedy finished on its own
    remedy_result(r, T, (f & AM_F_REMEDY_OUTCOME_OK) != 0, res);
  }
  r.flags = f & ~(AM_F_PENDING_OK | AM_F_PENDING_FAIL | AM_F_REMEDY_PENDING | AM_F_REMEDY_OUTCOME_OK);
  return act;
}

// ---------------------------------------------------------------------------
// The sweep.  CLOSED = closed-loop harness: a due record
// completes in the same tick with its preset outcome.
// ---------------------------------------------------------------------------
#ifndef AM_MIN_BLOCKS
#define AM_MIN_BLOCKS 3
#endif
// MASKS = the five cron-field columns are read and matched.  Off the minute a
// 5-field schedule cannot fire (ParseStandard pins Second to 1<<0), so the
// launcher picks MASKS=false unless AM_SWEEP_FULL_SCAN is set: 40 of the 56
// bytes per record and all of the mask arithmetic disappear at compile time.
template <bool CLOSED, bool MASKS>
__global__ void __launch_bounds__(kBlock, MASKS ? AM_MIN_BLOCKS : AM_MIN_BLOCKS + 1) sweep_tick_kernel(const SweepParams p) {
  // one row per warp, written unconditionally: no zero-initialisation, no shared
  // atomics, and therefore a single __syncthreads in the whole kernel
  __shared__ uint32_t s_warp_tot[kWarps];
  __shared__ uint32_t s_wres[kWarps][4];  // posted results applied: ok, fail, remedy ok, remedy fail

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const uint32_t tile = blockIdx.x;
  const uint32_t tile_base = tile * (uint32_t)kTile;
  const int64_t T = p.T;

  // ---- phase A: issue every schedule-column load of this lane up front ----
  // half h covers records r0(h) .. r0(h)+1, r0 = tile_base + warp*128 + h*64 + lane*2
  uint32_t r0[2];
  ulonglong2 mi[2], hr[2], dm[2], mo[2], dw[2];
  longlong2 fa[2];
  int2 ras[2];
  uint2 fl[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    r0[h] = tile_base + (uint32_t)(warp * kRecPerWarp + h * 64 + lane * 2);
    fl[h] = ld_stream(reinterpret_cast<const uint2*>(p.c.flags + r0[h]));
    ras[h] = ld_stream(reinterpret_cast<const int2*>(p.c.ras + r0[h]));
    fa[h] = ld_stream(reinterpret_cast<const longlong2*>(p.c.finished_at + r0[h]));
  }
  if (MASKS) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mi[h] = ld_stream(reinterpret_cast<const ulonglong2*>(p.c.minute + r0[h]));
      hr[h] = ld_stream(reinterpret_cast<const ulonglong2*>(p.c.hour + r0[h]));
      dm[h] = ld_stream(reinterpret_cast<const ulonglong2*>(p.c.dom + r0[h]));
      mo[h] = ld_stream(reinterpret_cast<const ulonglong2*>(p.c.month + r0[h]));
      dw[h] = ld_stream(reinterpret_cast<const ulonglong2*>(p.c.dow + r0[h]));
    }
  }

  // Scheduling fence: a warp-level memory-ordering point.  ptxas may not sink the loads
  // above across it, so all 16 are issued before any of the arithmetic below (without it
  // the second half's loads were delayed behind the first half's compute: -30 % bandwidth).
  __syncwarp();

  // The tick's broken-down time: one-hot words computed once per tick (civil.h) and
  // delivered through the kernel parameters, i.e. the constant bank / uniform
  // registers — cheaper than staging them in shared memory, which cost every CTA a
  // serial thread-0 section and a barrier (profiles/r01_summary.md).
  const TickWords w = p.words;

  uint32_t act[2][2];
  uint32_t res_lane = 0;  // 4 x 8-bit counts of results applied by this lane
  uint32_t nfl[2][2];     // flags / finishedAt after this tick
  int64_t nfa[2][2];
  bool dirty[2] = {false, false};
  uint32_t needy = 0, due_bits = 0;  // bit (2h+j): record needs the remedy/counter columns / is due

  // ---- schedule decision for the lane's four records ------------------------
#pragma unroll
  for (int h = 0; h < 2; ++h) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint32_t f = j ? fl[h].y : fl[h].x;
      const int32_t rasv = j ? ras[h].y : ras[h].x;
      const int64_t fav = j ? fa[h].y : fa[h].x;
      const uint32_t kind = f & AM_KIND_MASK;
      // kinds 1..5 are evaluated; tombstones, NO_RESOURCE (hcc.go:227) and host-fallback are not
      const bool live = ((0x3Eu >> kind) & 1u) && !(f & AM_F_TOMBSTONE);
      const bool has_result = (f & (AM_F_PENDING_OK | AM_F_PENDING_FAIL)) != 0;
      const bool pending = (f & (AM_F_PENDING_OK | AM_F_PENDING_FAIL | AM_F_REMEDY_PENDING)) != 0;
      // step 1 sets finishedAt = T before the due decision is taken
      const int64_t fa_eff = has_result ? T : fav;
      const int64_t elapsed = (int64_t)((uint64_t)T - (uint64_t)fa_eff);
      const bool due_iv = !(elapsed < (int64_t)rasv);  // not(hcc.go:264) == timer :751 fired
      bool due_cron = false;
      if (MASKS) {
        const uint64_t miv = j ? mi[h].y : mi[h].x, hrv = j ? hr[h].y : hr[h].x;
        const uint64_t dmv = j ? dm[h].y : dm[h].x, mov = j ? mo[h].y : mo[h].x;
        const uint64_t dwv = j ? dw[h].y : dw[h].x;
        // branch-free: every term is evaluated (no short-circuit control flow)
        const bool fld = ((miv & w.minute) != 0) & ((hrv & w.hour) != 0) & ((mov & w.month) != 0);
        const bool dmm = (dmv & w.dom) != 0, dwm = (dwv & w.dow) != 0;
        const bool star = ((dmv | dwv) >> 63) != 0;  // robfig dayMatches
        due_cron = (w.sec0 != 0) & fld & (star ? (dmm & dwm) : (dmm | dwm));
      }
      const bool is_iv = ((0x14u >> kind) & 1u) != 0;  // INTERVAL or CRON_EVERY
      const bool due = live && (is_iv ? due_iv : (kind == AM_KIND_CRON_SPEC && due_cron));
      const bool stopped_now = live && kind == AM_KIND_STOPPED && !(f & AM_F_STOPPED_REPORTED);
      act[h][j] = (due ? AM_ACT_SUBMIT_HC : 0u) | (stopped_now ? AM_ACT_STOPPED : 0u) |
                  ((live && kind == AM_KIND_PARSE_ERROR) ? AM_ACT_PARSE_ERROR : 0u);
      nfl[h][j] = f;
      nfa[h][j] = fav;
      if (stopped_now) {  // hcc.go:238-250: Status "Stopped", FinishedAt = now
        nfl[h][j] = f | AM_F_STOPPED_REPORTED;
        nfa[h][j] = T;
        dirty[h] = true;
      }
      if (live && (pending || (CLOSED && due))) needy |= 1u << (2 * h + j);
      if (due) due_bits |= 1u << (2 * h + j);
    }
  }

  // ---- results + remedy state machine (hcc.go:633-724, 819-852) -------------------
  // Two shapes, chosen per warp by how many lanes have work:
  //  dense  (>= 16 lanes, e.g. half of all checks reporting a result this tick): per
  //         pair of records, vector loads of the eight remedy/counter columns and the
  //         state machine predicated per record — the bytes are needed anyway and 16-B
  //         transactions keep the LSU count low;
  //  sparse (a few due/posted records per warp): a warp loop in which every lane takes
  //         its next needy record with scalar accesses — the loop runs once or twice
  //         instead of four predicated copies of the state machine.
  if (__popc(__ballot_sync(kFull, needy != 0)) >= 16) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint32_t nb = (needy >> (2 * h)) & 3u;
      if (nb) {
        const uint32_t r = r0[h];
        const int2 lim = ld_stream(reinterpret_cast<const int2*>(p.c.runs_limit + r));
        const int2 rst = ld_stream(reinterpret_cast<const int2*>(p.c.reset_interval + r));
        const int2 sc = ld_stream(reinterpret_cast<const int2*>(p.c.success + r));
        const int2 fc = ld_stream(reinterpret_cast<const int2*>(p.c.failed + r));
        const int2 rsc = ld_stream(reinterpret_cast<const int2*>(p.c.remedy_success + r));
        const int2 rfc = ld_stream(reinterpret_cast<const int2*>(p.c.remedy_failed + r));
        const int2 rtc = ld_stream(reinterpret_cast<const int2*>(p.c.remedy_total + r));
        const longlong2 rfa = ld_stream(reinterpret_cast<const longlong2*>(p.c.remedy_finished_at + r));
        int32_t ns[2] = {sc.x, sc.y}, nf[2] = {fc.x, fc.y};
        int32_t nrs[2] = {rsc.x, rsc.y}, nrf[2] = {rfc.x, rfc.y}, nrt[2] = {rtc.x, rtc.y};
        int64_t nrfa[2] = {rfa.x, rfa.y};
        const int32_t limv[2] = {lim.x, lim.y}, rstv[2] = {rst.x, rst.y};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if ((nb >> j) & 1u) {
            RecState s{nfl[h][j], nfa[h][j], ns[j], nf[j], nrs[j], nrf[j], nrt[j], nrfa[j], limv[j], rstv[j]};
            uint32_t res = 0;
            uint32_t a = apply_result(s, T, res);
            if (CLOSED && ((due_bits >> (2 * h + j)) & 1u)) {
              const uint64_t k = outcome_key(p.seed, p.shard_base + r + (uint32_t)j, (uint64_t)T);
              const uint32_t failp = (s.flags >> AM_F_FAILP_SHIFT) & 0xFFu;
              const bool fail = (uint32_t)(k & 0xFF) < failp;
              const bool rem_ok = (uint32_t)((k >> 8) & 0xFF) < 179u;
              s.flags |= (fail ? AM_F_PENDING_FAIL : AM_F_PENDING_OK) | AM_F_REMEDY_PENDING |
                         (rem_ok ? AM_F_REMEDY_OUTCOME_OK : 0u);
              a |= apply_result(s, T, res);
            }
            act[h][j] |= a;
            res_lane += res;
            nfl[h][j] = s.flags; nfa[h][j] = s.fa;
            ns[j] = s.s; nf[j] = s.f; nrs[j] = s.rs; nrf[j] = s.rf; nrt[j] = s.rt; nrfa[j] = s.rfa;
          }
        }
        dirty[h] = true;  // a result always clears its PENDING flags
        if (ns[0] != sc.x || ns[1] != sc.y) st_stream(reinterpret_cast<int2*>(p.c.success + r), make_int2(ns[0], ns[1]));
        if (nf[0] != fc.x || nf[1] != fc.y) st_stream(reinterpret_cast<int2*>(p.c.failed + r), make_int2(nf[0], nf[1]));
        if (nrs[0] != rsc.x || nrs[1] != rsc.y)
          st_stream(reinterpret_cast<int2*>(p.c.remedy_success + r), make_int2(nrs[0], nrs[1]));
        if (nrf[0] != rfc.x || nrf[1] != rfc.y)
          st_stream(reinterpret_cast<int2*>(p.c.remedy_failed + r), make_int2(nrf[0], nrf[1]));
        if (nrt[0] != rtc.x || nrt[1] != rtc.y)
          st_stream(reinterpret_cast<int2*>(p.c.remedy_total + r), make_int2(nrt[0], nrt[1]));
        if (nrfa[0] != rfa.x || nrfa[1] != rfa.y)
          st_stream(reinterpret_cast<longlong2*>(p.c.remedy_finished_at + r), make_longlong2(nrfa[0], nrfa[1]));
      }
    }
    needy = 0;
  }
  while (__any_sync(kFull, needy != 0)) {
    if (needy) {
      const int b = __ffs(needy) - 1;
      needy &= needy - 1;
      const uint32_t i = r0[0] + (uint32_t)(64 * (b >> 1) + (b & 1));
      const int32_t lim = ld_stream(p.c.runs_limit + i), rst = ld_stream(p.c.reset_interval + i);
      const int32_t sc = ld_stream(p.c.success + i), fc = ld_stream(p.c.failed + i);
      const int32_t rsc = ld_stream(p.c.remedy_success + i), rfc = ld_stream(p.c.remedy_failed + i);
      const int32_t rtc = ld_stream(p.c.remedy_total + i);
      const int64_t rfa = ld_stream(p.c.remedy_finished_at + i);
      const uint32_t f0 = b == 0 ? nfl[0][0] : b == 1 ? nfl[0][1] : b == 2 ? nfl[1][0] : nfl[1][1];
      const int64_t fa0 = b == 0 ? nfa[0][0] : b == 1 ? nfa[0][1] : b == 2 ? nfa[1][0] : nfa[1][1];
      RecState s{f0, fa0, sc, fc, rsc, rfc, rtc, rfa, lim, rst};
      uint32_t res = 0;
      uint32_t a = apply_result(s, T, res);
      if (CLOSED && ((due_bits >> b) & 1u)) {
        const uint64_t k = outcome_key(p.seed, p.shard_base + i, (uint64_t)T);
        const uint32_t failp = (s.flags >> AM_F_FAILP_SHIFT) & 0xFFu;
        const bool fail = (uint32_t)(k & 0xFF) < failp;
        const bool rem_ok = (uint32_t)((k >> 8) & 0xFF) < 179u;
        s.flags |= (fail ? AM_F_PENDING_FAIL : AM_F_PENDING_OK) | AM_F_REMEDY_PENDING |
                   (rem_ok ? AM_F_REMEDY_OUTCOME_OK : 0u);
        a |= apply_result(s, T, res);
      }
      res_lane += res;  // per lane at most 4 records x 2 results per byte
      if (s.s != sc) st_stream(p.c.success + i, s.s);
      if (s.f != fc) st_stream(p.c.failed + i, s.f);
      if (s.rs != rsc) st_stream(p.c.remedy_success + i, s.rs);
      if (s.rf != rfc) st_stream(p.c.remedy_failed + i, s.rf);
      if (s.rt != rtc) st_stream(p.c.remedy_total + i, s.rt);
      if (s.rfa != rfa) st_stream(p.c.remedy_finished_at + i, s.rfa);
      // (a record "Stopped" in this very tick already carries STOPPED_REPORTED and
      // finishedAt = T in nfl/nfa, and apply_result preserves both: the pause rule of
      // hcc.go:238-250 and a posted result commute)
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q == b) {
          act[q >> 1][q & 1] |= a;
          nfl[q >> 1][q & 1] = s.flags;
          nfa[q >> 1][q & 1] = s.fa;
        }
      }
      if (b < 2) dirty[0] = true; else dirty[1] = true;  // a result always clears its PENDING flags
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    if (dirty[h]) {
      st_stream(reinterpret_cast<uint2*>(p.c.flags + r0[h]), make_uint2(nfl[h][0], nfl[h][1]));
      st_stream(reinterpret_cast<longlong2*>(p.c.finished_at + r0[h]), make_longlong2(nfa[h][0], nfa[h][1]));
    }
  }

  // ---- ordered compaction: in-warp ranks from ballots ---------------------
  const unsigned lt = (1u << lane) - 1u;
  const unsigned b00 = __ballot_sync(kFull, act[0][0] != 0), b01 = __ballot_sync(kFull, act[0][1] != 0);
  const unsigned b10 = __ballot_sync(kFull, act[1][0] != 0), b11 = __ballot_sync(kFull, act[1][1] != 0);
  const uint32_t tot0 = __popc(b00) + __popc(b01);
  const uint32_t warp_total = tot0 + __popc(b10) + __popc(b11);
  uint32_t rank[2][2];
  rank[0][0] = __popc(b00 & lt) + __popc(b01 & lt);
  rank[0][1] = rank[0][0] + (act[0][0] != 0);
  rank[1][0] = tot0 + __popc(b10 & lt) + __popc(b11 & lt);
  rank[1][1] = rank[1][0] + (act[1][0] != 0);
  if (lane == 0) s_warp_tot[warp] = warp_total;

  // ---- results applied this tick (feeds metrics.MonitorSuccess/Error): lane ->
  //      warp (two redux over 16-bit halves) -> per-warp shared row.  The action
  //      statistics and index checksums are derived from the emitted entries
  //      by compact_kernel, so records that emit nothing cost nothing here.
  {
    uint32_t lo = 0, hi = 0;
    if (__any_sync(kFull, res_lane != 0)) {
      lo = __reduce_add_sync(kFull, (res_lane & 0xFFu) | ((res_lane & 0xFF00u) << 8));           // ok | fail<<16
      hi = __reduce_add_sync(kFull, ((res_lane >> 16) & 0xFFu) | ((res_lane >> 8) & 0xFF0000u));  // remedy ok | fail<<16
    }
    if (lane < 4) s_wres[warp][lane] = ((lane < 2 ? lo : hi) >> ((lane & 1) * 16)) & 0xFFFFu;
  }
  __syncthreads();

  // ---- in-tile base, segment write-out, per-tile count ----------------------
  const uint64_t keep = l2_evict_last_policy();
  uint32_t base = tile_base, tile_total = 0;
#pragma unroll
  for (int k = 0; k < kWarps; ++k) {
    const uint32_t v = s_warp_tot[k];
    base += (k < warp) ? v : 0u;
    tile_total += v;
  }
#pragma unroll
  for (int h = 0; h < 2; ++h)
#pragma unroll
    for (int j = 0; j < 2; ++j)
      if (act[h][j]) {
        const uint32_t pos = base + rank[h][j];
        st_keep_u32(p.seg_idx + pos, r0[h] + (uint32_t)j, keep);
        st_keep_u8(p.seg_act + pos, act[h][j], keep);
      }
  if (warp == 0) {  // per-tile count, group counter, result counters (RED, no return value)
    if (lane < 4) {
      uint32_t sv = 0;
#pragma unroll
      for (int k = 0; k < kWarps; ++k) sv += s_wres[k][lane];
      if (sv) atomicAdd(&p.acc[10 + lane], (unsigned long long)sv);
    }
    if (lane == 4) {
      p.tile_count[tile] = tile_total;
      if (tile_total) atomicAdd(&p.group_count[tile / kGroupTiles], tile_total);
    }
  }
}

// ---------------------------------------------------------------------------
#ifndef AM_COMPACT_UNROLL
#define AM_COMPACT_UNROLL 4  // independent (index, action) loads in flight per thread; build-time knob
#endif
constexpr int kCompactUnroll = AM_COMPACT_UNROLL;

// Segments -> contiguous ascending list.  One CTA per group of kGroupTiles
// tiles: its global base is the sum of the earlier groups' counts (a few KB of
// L2-resident reads), the in-group offsets a kGroupTiles-wide scan; entries just written
// by the sweep are still in L2.  While moving its entries every thread also counts their
// action bits and checksums their indices (statistics cost is proportional to what
// was emitted, not to N); every group zeroes its slot of the other-parity group
// counters for the next tick.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) compact_kernel(const CompactParams p) {
  __shared__ uint32_t s_part[8];
  __shared__ uint32_t s_off[kGroupTiles + 1];
  __shared__ uint32_t s_cnt[8][8];               // per-warp counts of the 8 action bits
  __shared__ unsigned long long s_chk[8][2];     // per-warp xor / sum of emitted global indices
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t g = blockIdx.x;

  uint32_t part = 0;
  for (uint32_t j = tid; j < g; j += blockDim.x) part += p.group_count[j];
  part = __reduce_add_sync(kFull, part);
  if (lane == 0) s_part[warp] = part;
  if (warp == 0) {  // exclusive scan of this group's tile counts (first kGroupTiles lanes)
    const uint32_t t = g * kGroupTiles + (uint32_t)lane;
    const uint32_t c = (lane < kGroupTiles && t < p.n_tiles) ? p.tile_count[t] : 0u;
    uint32_t incl = c;
#pragma unroll
    for (int d = 1; d < kGroupTiles; d <<= 1) {
      const uint32_t up = __shfl_up_sync(kFull, incl, d);
      if (lane >= d) incl += up;
    }
    if (lane < kGroupTiles) s_off[lane] = incl - c;
    if (lane == kGroupTiles - 1) s_off[kGroupTiles] = incl;
  }
  __syncthreads();
  uint32_t base = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k) base += s_part[k];
  const uint32_t group_total = s_off[kGroupTiles];

  // per-thread statistics of the entries it moves: 8 action-bit counts (two words of
  // four bytes; a thread moves kGroupTiles*kTile/256 <= 128 entries) and the checksums of the global indices
  uint32_t c0 = 0, c1 = 0;
  unsigned long long cx = 0, cs = 0;
  // kCompactUnroll independent (index, action) loads in flight per thread
  for (uint32_t e0 = tid; e0 < group_total; e0 += (uint32_t)kCompactUnroll * blockDim.x) {
    uint32_t src[kCompactUnroll], vi[kCompactUnroll], va[kCompactUnroll];
#pragma unroll
    for (int u = 0; u < kCompactUnroll; ++u) {
      const uint32_t e = e0 + (uint32_t)u * blockDim.x;
      int tt = 0;  // which tile of the group holds entry e
#pragma unroll
      for (int k = 1; k < kGroupTiles; ++k) tt += (e >= s_off[k]) ? 1 : 0;
      src[u] = (g * kGroupTiles + (uint32_t)tt) * (uint32_t)kTile + (e - s_off[tt]);
    }
#pragma unroll
    for (int u = 0; u < kCompactUnroll; ++u) {
      const bool ok = e0 + (uint32_t)u * blockDim.x < group_total;
      vi[u] = ok ? __ldcs(p.seg_idx + src[u]) : 0u;
      va[u] = ok ? (uint32_t)__ldcs(p.seg_act + src[u]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kCompactUnroll; ++u) {
      const uint32_t e = e0 + (uint32_t)u * blockDim.x;
      const uint32_t pos = base + e;
      if (e < group_total) {
        c0 += spread4(va[u]);
        c1 += spread4(va[u] >> 4);
        const unsigned long long gi = p.shard_base + vi[u];
        cx ^= gi;
        cs += gi;
        if (pos < p.cap) {
          p.out_idx[pos] = vi[u];
          p.out_act[pos] = (uint8_t)va[u];
        }
      }
    }
  }
  // statistics: thread -> warp (redux; 16-bit fields so 32 lanes x 63 fit) -> CTA -> global RED
  if (group_total) {
    const uint32_t a0 = __reduce_add_sync(kFull, (c0 & 0xFFu) | ((c0 & 0xFF00u) << 8));
    const uint32_t a1 = __reduce_add_sync(kFull, ((c0 >> 16) & 0xFFu) | ((c0 >> 8) & 0xFF0000u));
    const uint32_t a2 = __reduce_add_sync(kFull, (c1 & 0xFFu) | ((c1 & 0xFF00u) << 8));
    const uint32_t a3 = __reduce_add_sync(kFull, ((c1 >> 16) & 0xFFu) | ((c1 >> 8) & 0xFF0000u));
    const uint32_t xl = __reduce_xor_sync(kFull, (uint32_t)cx), xh = __reduce_xor_sync(kFull, (uint32_t)(cx >> 32));
    unsigned long long sum = cs;
#pragma unroll
    for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(kFull, sum, d);
    if (lane < 8) {
      const uint32_t q = lane < 2 ? a0 : (lane < 4 ? a1 : (lane < 6 ? a2 : a3));
      s_cnt[warp][lane] = (q >> ((lane & 1) * 16)) & 0xFFFFu;
    }
    if (lane == 0) { s_chk[warp][0] = ((unsigned long long)xh << 32) | xl; s_chk[warp][1] = sum; }
    __syncthreads();
    if (tid < 8) {
      uint32_t v = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) v += s_cnt[k][tid];
      if (v) atomicAdd(&p.acc[2 + tid], (unsigned long long)v);
    }
    if (tid == 8) {
      unsigned long long x = 0, t = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) { x ^= s_chk[k][0]; t += s_chk[k][1]; }
      atomicXor(&p.acc[14], x);
      atomicAdd(&p.acc[15], t);
    }
  }
  if (tid == 0) p.group_count_next[g] = 0;

  if (g == p.n_groups - 1 && tid == 0) p.acc[1] = (unsigned long long)base + group_total;  // n_emitted
}

// One warp, after the kernel boundary that completes every group's REDs: publish the
// tick's am_tick_stats_t (device or mapped-host memory) and re-arm the accumulators.
__global__ void publish_kernel(unsigned long long* acc, am_tick_stats_t* out_stats, uint32_t* out_count,
                               uint64_t n_records) {
  const int k = threadIdx.x;
  if (k >= kNumAcc) return;
  unsigned long long v = acc[k];
  acc[k] = 0;
  if (k == 0) v = n_records;
  if (out_stats) reinterpret_cast<unsigned long long*>(out_stats)[k] = v;
  if (k == 1 && out_count) *out_count = (uint32_t)v;
}

// ---- small maintenance kernels (create / read) -----------------------------
__global__ void fill_u32_kernel(uint32_t* p, uint32_t v, uint64_t n) {
  uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = v;
}

// RepeatAfterSec as the reference derives it at reconcile time (hcc.go:259-262):
// for a 5-field schedule Next(T) - T (robfig SpecSchedule.Next, cron_next_utc), for
// "@every d" and interval checks the stored interval, 0 for everything else.  One
// thread per record; the day scan diverges, which is fine for an on-demand query
// (status display, timer-wheel bucketing) that is not part of the per-tick path.
__global__ void next_fire_kernel(DevCols c, uint32_t first, uint32_t n, int64_t T, int64_t* out) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t i = first + k;
  const uint32_t f = c.flags[i];
  const uint32_t kind = f & AM_KIND_MASK;
  int64_t v = 0;
  if (!(f & AM_F_TOMBSTONE)) {
    if (kind == AM_KIND_CRON_SPEC)
      v = repeat_after_from_next(cron_next_utc(c.minute[i], c.hour[i], c.dom[i], c.month[i], c.dow[i], T), T);
    else if (kind == AM_KIND_INTERVAL || kind == AM_KIND_CRON_EVERY)
      v = c.ras[i];
  }
  out[k] = v;
}

// ---- staged controller events (upsert / remove / post_result) ---------------
// Events reach the library from many goroutines between two ticks; per slot
// they must take effect in call order.  An event's tick-local sequence number
// is its position in the staged array (1-based).  mark: atomicMax of the sequence into the slot's
// mark pair {latest upsert/remove, latest result}; apply: only the marked
// winners write — the latest upsert/remove, then the latest result if it was
// posted after it (an older result belongs to the replaced CR); clear: marks
// back to zero.  No host-side hashing or sorting.
struct StagedOp {
  uint32_t idx;  // local slot
  uint32_t arg;  // kind in the top 2 bits; low 30 bits: record index (upsert) or flag bits (result)
};               // the sequence number of an op is its position in the array + 1
constexpr uint32_t kOpUpsert = 0u << 30, kOpRemove = 1u << 30, kOpResult = 2u << 30, kOpKindMask = 3u << 30;

__global__ void mark_ops_kernel(uint32_t* marks, const StagedOp* __restrict__ ops, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const StagedOp op = ops[k];
  atomicMax(&marks[2u * op.idx + ((op.arg & kOpKindMask) == kOpResult ? 1u : 0u)], k + 1u);
}

// upserts (hcc.go:170-188 Reconcile) and removes (hcc.go:175-186)
__global__ void apply_state_ops_kernel(DevCols c, const uint32_t* __restrict__ marks,
                                       const StagedOp* __restrict__ ops,
                                       const am_record_t* __restrict__ recs, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const StagedOp op = ops[k];
  const uint32_t kind = op.arg & kOpKindMask;
  if (kind == kOpResult || marks[2u * op.idx] != k + 1u) return;
  const uint32_t i = op.idx;
  if (kind == kOpRemove) { c.flags[i] = AM_F_TOMBSTONE; return; }
  const am_record_t r = recs[op.arg & ~kOpKindMask];
  c.minute[i] = r.minute; c.hour[i] = r.hour; c.dom[i] = r.dom; c.month[i] = r.month; c.dow[i] = r.dow;
  c.ras[i] = r.ras; c.flags[i] = r.flags; c.finished_at[i] = r.finished_at;
  c.runs_limit[i] = r.runs_limit; c.reset_interval[i] = r.reset_interval;
  c.success[i] = r.success; c.failed[i] = r.failed; c.remedy_success[i] = r.remedy_success;
  c.remedy_failed[i] = r.remedy_failed; c.remedy_total[i] = r.remedy_total;
  c.remedy_finished_at[i] = r.remedy_finished_at;
}

// terminal phases observed by the watch loops (hcc.go:635/:662/:821/:836)
__global__ void apply_result_ops_kernel(uint32_t* flags, const uint32_t* __restrict__ marks,
                                        const StagedOp* __restrict__ ops, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const StagedOp op = ops[k];
  if ((op.arg & kOpKindMask) != kOpResult) return;
  const uint32_t s = marks[2u * op.idx], r = marks[2u * op.idx + 1u];
  if (r != k + 1u || k + 1u < s) return;
  const uint32_t m = AM_F_PENDING_OK | AM_F_PENDING_FAIL | AM_F_REMEDY_PENDING | AM_F_REMEDY_OUTCOME_OK;
  flags[op.idx] = (flags[op.idx] & ~m) | (op.arg & m);
}

__global__ void clear_marks_kernel(uint32_t* marks, const StagedOp* __restrict__ ops, uint32_t n) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t i = ops[k].idx;
  marks[2u * i] = 0;
  marks[2u * i + 1u] = 0;
}

__global__ void gather_records_kernel(DevCols c, const uint32_t* __restrict__ idx, am_record_t* out,
                                      uint32_t n) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const uint32_t i = idx[k];
  am_record_t r;
  r.minute = c.minute[i]; r.hour = c.hour[i]; r.dom = c.dom[i]; r.month = c.month[i]; r.dow = c.dow[i];
  r.finished_at = c.finished_at[i]; r.remedy_finished_at = c.remedy_finished_at[i];
  r.ras = c.ras[i]; r.flags = c.flags[i];
  r.runs_limit = c.runs_limit[i]; r.reset_interval = c.reset_interval[i];
  r.success = c.success[i]; r.failed = c.failed[i]; r.remedy_success = c.remedy_success[i];
  r.remedy_failed = c.remedy_failed[i]; r.remedy_total = c.remedy_total[i];
  r.reserved = 0;
  out[k] = r;
}

}  // namespace amsweep
