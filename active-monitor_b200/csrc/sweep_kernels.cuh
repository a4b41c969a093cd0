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
//     1024 consecutive records, a warp owns 128 consecutive ones;
//   * the tile's schedule columns are staged in shared memory by TMA bulk copies
//     (cp.async.bulk + mbarrier): one elected thread issues one 4-8 KB copy per
//     column, 56 KB in flight per CTA independent of register allocation, no
//     per-lane load instructions; thread t then owns records 4t..4t+3;
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
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/amsweep.h"
#include "civil.h"

namespace amsweep {

#ifndef AM_BLOCK
#define AM_BLOCK 256
#endif
constexpr int kBlock = AM_BLOCK;
constexpr int kWarps = kBlock / 32;
constexpr int kRecPerWarp = 128;             // 32 lanes x 4 consecutive records
constexpr int kTile = kWarps * kRecPerWarp;  // 1024 records per CTA
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kNumAcc = 16;  // == number of u64 fields of am_tick_stats_t
// phase-A staging per CTA (one tile in record order): finishedAt + ras + flags [+ 5 masks]
constexpr size_t kStageBytesNoMasks = (size_t)kTile * (8 + 4 + 4);           // 16 KB
constexpr size_t kStageBytesMasks = kStageBytesNoMasks + (size_t)kTile * 40;  // 56 KB

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

constexpr int kGroupTiles = 8;  // tiles per compaction group

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

// Phase-A staging: TMA bulk copies (cp.async.bulk, SASS UBLKCP) global -> shared,
// completion counted in bytes on an mbarrier.  One elected thread issues one copy per
// column (4-8 KB each) for the whole tile; no lane executes a load for phase A, the
// memory-level parallelism (56 KB per CTA) does not depend on register allocation, and
// the data sits in shared memory in record order.  Measured alternatives: register
// loads were split into two batches by ptxas and once even sunk into the match
// branches (-45 % bandwidth); per-lane cp.async (LDGSTS) cost 16 copy + 16 read
// instructions per lane (profiles/r01_summary.md).
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"((uint32_t)__cvta_generic_to_shared(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"((uint32_t)__cvta_generic_to_shared(bar)),
      "r"(phase)
      : "memory");
}

template <typename T>
__device__ __forceinline__ void st_stream(T* p, T v) { __stcs(p, v); }

// Segment entries are written once by the sweep and read once, a few tens of
// microseconds later, by compact_kernel: keep them L2-resident (evict_last)
// while 560 MB of evict_first column data streams past them.
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
  // one row per warp, written unconditionally: no zero-initialisation, no shared atomics
  __shared__ uint32_t s_warp_tot[kWarps];
  __shared__ uint32_t s_wres[kWarps][4];  // posted results applied: ok, fail, remedy ok, remedy fail

  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int warp = tid >> 5;
  const int64_t T = p.T;
  // The tick's broken-down time: one-hot words computed once per tick (civil.h) and
  // delivered through the kernel parameters, i.e. the constant bank / uniform registers.
  const TickWords w = p.words;
  const uint64_t keep = l2_evict_last_policy();

  // ---- phase A: a two-stage TMA pipeline over this CTA's tiles -------------------
  // The CTA is persistent: it owns tiles blockIdx.x, blockIdx.x + gridDim.x, ...  While
  // tile i is evaluated out of one shared-memory stage, the bulk copies of tile i+1 are
  // already in flight into the other, so HBM stays busy through the compute, the
  // write-out and the barriers.  Stage layout (record order): [fa 8K][ras 4K][flags 4K]
  // then, with MASKS, [minute][hour][dom][month][dow] 8K each; thread t owns records
  // 4t..4t+3 of the tile.
  extern __shared__ __align__(128) unsigned char stage_mem[];
  __shared__ __align__(8) uint64_t s_bar[2];
  constexpr uint32_t kStageBytes = (uint32_t)(MASKS ? kStageBytesMasks : kStageBytesNoMasks);
  auto issue_tile = [&](uint32_t t, int st) {  // called by thread 0 only
    unsigned char* base = stage_mem + (size_t)st * kStageBytes;
    const uint32_t tb = t * (uint32_t)kTile;
    mbar_expect_tx(&s_bar[st], kStageBytes);
    tma_bulk_g2s(base, p.c.finished_at + tb, kTile * 8, &s_bar[st]);
    tma_bulk_g2s(base + kTile * 8, p.c.ras + tb, kTile * 4, &s_bar[st]);
    tma_bulk_g2s(base + kTile * 12, p.c.flags + tb, kTile * 4, &s_bar[st]);
    if (MASKS) {
      uint64_t* m = reinterpret_cast<uint64_t*>(base + kTile * 16);
      tma_bulk_g2s(m + 0 * kTile, p.c.minute + tb, kTile * 8, &s_bar[st]);
      tma_bulk_g2s(m + 1 * kTile, p.c.hour + tb, kTile * 8, &s_bar[st]);
      tma_bulk_g2s(m + 2 * kTile, p.c.dom + tb, kTile * 8, &s_bar[st]);
      tma_bulk_g2s(m + 3 * kTile, p.c.month + tb, kTile * 8, &s_bar[st]);
      tma_bulk_g2s(m + 4 * kTile, p.c.dow + tb, kTile * 8, &s_bar[st]);
    }
  };
  if (tid == 0) {
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    fence_mbar_init();
    if (blockIdx.x < p.n_tiles) issue_tile(blockIdx.x, 0);
  }
  __syncthreads();  // the initialised barriers are visible to every waiter

  uint32_t it = 0;
  for (uint32_t tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
  const int st = (int)(it & 1u);
  const uint32_t tile_base = tile * (uint32_t)kTile;
  // prefetch the next tile into the other stage (its previous readers passed the barrier
  // that ends the previous iteration)
  if (tid == 0 && tile + gridDim.x < p.n_tiles) issue_tile(tile + gridDim.x, st ^ 1);
  mbar_wait(&s_bar[st], (it >> 1) & 1u);
  const unsigned char* stage = stage_mem + (size_t)st * kStageBytes;
  const int64_t* s_fa = reinterpret_cast<const int64_t*>(stage);
  const int32_t* s_ras = reinterpret_cast<const int32_t*>(stage + kTile * 8);
  const uint32_t* s_flags = reinterpret_cast<const uint32_t*>(stage + kTile * 12);
  const uint64_t* s_mask = reinterpret_cast<const uint64_t*>(stage + kTile * 16);  // 5 x kTile

  const uint32_t r0 = tile_base + 4u * (uint32_t)tid;  // this thread's first record
  const uint4 fl4 = *reinterpret_cast<const uint4*>(s_flags + 4 * tid);
  const int4 ras4 = *reinterpret_cast<const int4*>(s_ras + 4 * tid);
  const longlong2 fa01 = *reinterpret_cast<const longlong2*>(s_fa + 4 * tid);
  const longlong2 fa23 = *reinterpret_cast<const longlong2*>(s_fa + 4 * tid + 2);
  const uint32_t flv[4] = {fl4.x, fl4.y, fl4.z, fl4.w};
  const int32_t rasv[4] = {ras4.x, ras4.y, ras4.z, ras4.w};
  const int64_t fav[4] = {fa01.x, fa01.y, fa23.x, fa23.y};

  uint32_t act[4];
  uint32_t res_lane = 0;  // 4 x 8-bit counts of results applied by this lane
  uint32_t nfl[4];        // flags / finishedAt after this tick
  int64_t nfa[4];
  bool dirty = false;
  uint32_t needy = 0, due_bits = 0;  // bit j: record needs the remedy/counter columns / is due

  // ---- schedule decision for the thread's four records ------------------------
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t f = flv[j];
    const uint32_t kind = f & AM_KIND_MASK;
    // kinds 1..5 are evaluated; tombstones, NO_RESOURCE (hcc.go:227) and host-fallback are not
    const bool live = ((0x3Eu >> kind) & 1u) && !(f & AM_F_TOMBSTONE);
    const bool has_result = (f & (AM_F_PENDING_OK | AM_F_PENDING_FAIL)) != 0;
    const bool pending = (f & (AM_F_PENDING_OK | AM_F_PENDING_FAIL | AM_F_REMEDY_PENDING)) != 0;
    // step 1 sets finishedAt = T before the due decision is taken
    const int64_t fa_eff = has_result ? T : fav[j];
    const int64_t elapsed = (int64_t)((uint64_t)T - (uint64_t)fa_eff);
    const bool due_iv = !(elapsed < (int64_t)rasv[j]);  // not(hcc.go:264) == timer :751 fired
    bool due_cron = false;
    if (MASKS) {
      const uint32_t k = 4u * (uint32_t)tid + (uint32_t)j;
      const uint64_t miv = s_mask[0 * kTile + k], hrv = s_mask[1 * kTile + k];
      const uint64_t dmv = s_mask[2 * kTile + k], mov = s_mask[3 * kTile + k];
      const uint64_t dwv = s_mask[4 * kTile + k];
      // branch-free: every term is evaluated (no short-circuit control flow)
      const bool fld = ((miv & w.minute) != 0) & ((hrv & w.hour) != 0) & ((mov & w.month) != 0);
      const bool dmm = (dmv & w.dom) != 0, dwm = (dwv & w.dow) != 0;
      const bool star = ((dmv | dwv) >> 63) != 0;  // robfig dayMatches
      due_cron = (w.sec0 != 0) & fld & (star ? (dmm & dwm) : (dmm | dwm));
    }
    const bool is_iv = ((0x14u >> kind) & 1u) != 0;  // INTERVAL or CRON_EVERY
    const bool due = live && (is_iv ? due_iv : (kind == AM_KIND_CRON_SPEC && due_cron));
    const bool stopped_now = live && kind == AM_KIND_STOPPED && !(f & AM_F_STOPPED_REPORTED);
    act[j] = (due ? AM_ACT_SUBMIT_HC : 0u) | (stopped_now ? AM_ACT_STOPPED : 0u) |
             ((live && kind == AM_KIND_PARSE_ERROR) ? AM_ACT_PARSE_ERROR : 0u);
    nfl[j] = f;
    nfa[j] = fav[j];
    if (stopped_now) {  // hcc.go:238-250: Status "Stopped", FinishedAt = now
      nfl[j] = f | AM_F_STOPPED_REPORTED;
      nfa[j] = T;
      dirty = true;
    }
    if (live && (pending || (CLOSED && due))) needy |= 1u << j;
    if (due) due_bits |= 1u << j;
  }

  // ---- results + remedy state machine: a warp loop in which every lane takes its
  //      next needy record.  With few posted results / due records per warp the loop
  //      runs once or twice instead of four predicated copies of the state machine;
  //      the remedy/counter columns are read and written per record (the 36 B/record
  //      are only touched for records that need them).
  while (__any_sync(kFull, needy != 0)) {
    if (needy) {
      const int b = __ffs(needy) - 1;
      needy &= needy - 1;
      const uint32_t i = r0 + (uint32_t)b;
      const int32_t lim = ld_stream(p.c.runs_limit + i), rst = ld_stream(p.c.reset_interval + i);
      const int32_t sc = ld_stream(p.c.success + i), fc = ld_stream(p.c.failed + i);
      const int32_t rsc = ld_stream(p.c.remedy_success + i), rfc = ld_stream(p.c.remedy_failed + i);
      const int32_t rtc = ld_stream(p.c.remedy_total + i);
      const int64_t rfa = ld_stream(p.c.remedy_finished_at + i);
      // (a record "Stopped" in this very tick already carries STOPPED_REPORTED and
      // finishedAt = T in nfl/nfa, and apply_result preserves both: the pause rule of
      // hcc.go:238-250 and a posted result commute)
      const uint32_t f0 = b == 0 ? nfl[0] : b == 1 ? nfl[1] : b == 2 ? nfl[2] : nfl[3];
      const int64_t fa0 = b == 0 ? nfa[0] : b == 1 ? nfa[1] : b == 2 ? nfa[2] : nfa[3];
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
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (q == b) {
          act[q] |= a;
          nfl[q] = s.flags;
          nfa[q] = s.fa;
        }
      }
      dirty = true;  // a result always clears its PENDING flags
    }
  }
  if (dirty) {
    st_stream(reinterpret_cast<uint4*>(p.c.flags + r0), make_uint4(nfl[0], nfl[1], nfl[2], nfl[3]));
    st_stream(reinterpret_cast<longlong2*>(p.c.finished_at + r0), make_longlong2(nfa[0], nfa[1]));
    st_stream(reinterpret_cast<longlong2*>(p.c.finished_at + r0 + 2), make_longlong2(nfa[2], nfa[3]));
  }

  // ---- ordered compaction: in-warp ranks from ballots (record order = lane, then j) ----
  const unsigned lt = (1u << lane) - 1u;
  uint32_t below = 0, warp_total = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned bj = __ballot_sync(kFull, act[j] != 0);
    below += __popc(bj & lt);
    warp_total += __popc(bj);
  }
  uint32_t rank[4];
  rank[0] = below;
  rank[1] = rank[0] + (act[0] != 0);
  rank[2] = rank[1] + (act[1] != 0);
  rank[3] = rank[2] + (act[2] != 0);
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
  uint32_t base = tile_base, tile_total = 0;
#pragma unroll
  for (int k = 0; k < kWarps; ++k) {
    const uint32_t v = s_warp_tot[k];
    base += (k < warp) ? v : 0u;
    tile_total += v;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (act[j]) {
      const uint32_t pos = base + rank[j];
      st_keep_u32(p.seg_idx + pos, r0 + (uint32_t)j, keep);
      st_keep_u8(p.seg_act + pos, act[j], keep);
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
  // everyone is done with this stage and with the per-warp rows before they are reused
  __syncthreads();
  }  // tile loop
}

// ---------------------------------------------------------------------------
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
  // four bytes; a thread moves < 64 entries) and the checksums of the global indices
  uint32_t c0 = 0, c1 = 0;
  unsigned long long cx = 0, cs = 0;
  // four independent (index, action) loads in flight per thread
  for (uint32_t e0 = tid; e0 < group_total; e0 += 4u * blockDim.x) {
    uint32_t src[4], vi[4], va[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t e = e0 + (uint32_t)u * blockDim.x;
      int tt = 0;  // which tile of the group holds entry e
#pragma unroll
      for (int k = 1; k < kGroupTiles; ++k) tt += (e >= s_off[k]) ? 1 : 0;
      src[u] = (g * kGroupTiles + (uint32_t)tt) * (uint32_t)kTile + (e - s_off[tt]);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const bool ok = e0 + (uint32_t)u * blockDim.x < group_total;
      vi[u] = ok ? __ldcs(p.seg_idx + src[u]) : 0u;
      va[u] = ok ? (uint32_t)__ldcs(p.seg_act + src[u]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
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
