// gather.cu — the global ascending due list on every GPU, over NVLink peer memory.
//
// The record array shards by contiguous global index range, so the global
// list is the rank-ordered concatenation of the shards' lists.  NCCL has no allgatherv; the
// portable path (gather.py) pads to the largest count and needs a host round-trip to size
// the exchange.  Here the exchange goes through CUDA-IPC-mapped peer memory, with no NCCL
// call and no host round-trip on the data path, in two forms:
//
//   am_gather_exchange (round 2, what bench.py uses)
//     ships the sweep's own per-tick output — 1 bit per record for the emitted set, the
//     group offsets, the non-default actions — into a per-rank slot of every peer's exchange
//     block (gather_push_tick_kernel: 16-B peer stores, system-scope done flags), then every
//     GPU rebuilds the GLOBAL (index, action) list from the world's bitmaps with the same
//     expand_kernel that rebuilds a single GPU's list.  ~1.3 MB per 10 M-record shard per
//     peer instead of 16.7 MB of finished entries.
//
//   am_gather_push (round 1 "plain" format, kept as the measured baseline)
//     every rank writes its finished (u32|u64 global index, u8 action) list straight into
//     every peer's output buffer at its global offset (counts exchanged in the same kernel).
//
// Slot and output buffers are double-buffered by epoch parity: a rank can be at most one
// epoch ahead of a peer, and each rank's consumers are stream-ordered before its next push,
// so epoch e+2 never overwrites data still being read.
//
// The reference has no counterpart (single Go process, no collectives); the consumer of
// the list is createSubmitWorkflow, hcc.go:502.
#ifndef AMSWEEP_EMULATE
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>

#include "gather_kernels.cuh"
#include "sweep_internal.h"

using namespace amsweep;

struct am_gather {
  int device = 0, rank = 0, world = 1;
  uint64_t cap_total = 0;
  int idx_bytes = 8;
  int n_ctas = 148;
  size_t block_bytes = 0;
  size_t off_idx[2] = {0, 0}, off_act[2] = {0, 0};
  unsigned char* block = nullptr;                 // my exchange block (cudaMalloc, IPC-exported)
  unsigned char* peer[kMaxWorld] = {};            // mapped peers (own = block)
  bool opened[kMaxWorld] = {};
  uint32_t* out_counts = nullptr;
  uint32_t* status = nullptr;                     // device word: 1 after a peer timed out
  uint32_t epoch = 0;
  bool connected = false;
  // tick exchange (am_gather_set_layout): per-rank slots for bitmap / prefix / tile_exc / exc_seg
  bool layout = false;
  size_t slots_off = 0, slots_bytes = 0;          // the slot area of the block (both parities)
  uint64_t bases[kMaxWorld] = {}, sizes[kMaxWorld] = {};
  uint32_t ngroups[kMaxWorld] = {}, ntiles[kMaxWorld] = {};
  size_t off_bitmap[2][kMaxWorld] = {}, off_prefix[2][kMaxWorld] = {}, off_tile_exc[2][kMaxWorld] = {},
         off_exc[2][kMaxWorld] = {};
  unsigned long long push_timeout_ms = 5000;      // AMSWEEP_PUSH_TIMEOUT_MS, 0 = none
  bool profiling = false, profiled = false;       // am_gather_set_profiling: events around the exchange's kernels
  cudaEvent_t evp[4] = {nullptr, nullptr, nullptr, nullptr};
  // am_gather_bind / am_gather_tick_view: one e2e step of a multi-GPU shard in one call
  am_sweep_t* bound = nullptr;
  cudaStream_t s_sweep = nullptr, s_exchange = nullptr;
  unsigned char* view_host = nullptr;   // mapped pinned: u32 idx[cap] | u8 act[cap] | u64 n | am_tick_stats_t
  uint64_t view_cap = 0;
  cudaEvent_t view_done = nullptr;
  std::string last_error;
};

#define AMG_CUDA(g, expr)                                                                       \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      char _b[384];                                                                             \
      snprintf(_b, sizeof _b, "%s -> %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      (g)->last_error = _b;                                                                     \
      return AM_E_DEVICE;                                                                       \
    }                                                                                           \
  } while (0)

namespace {
size_t align256(size_t v) { return (v + 255) / 256 * 256; }
}  // namespace

extern "C" {

int am_gather_create(am_gather_t** out, int device, int rank, int world, uint64_t cap_total,
                     int idx_bytes) {
  if (!out || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || cap_total == 0) return AM_E_INVAL;
  if (idx_bytes != 4 && idx_bytes != 8) return AM_E_INVAL;
  *out = nullptr;
  am_gather* g = new (std::nothrow) am_gather();
  if (!g) return AM_E_NOMEM;
  g->device = device; g->rank = rank; g->world = world; g->cap_total = cap_total;
  g->idx_bytes = idx_bytes;
  size_t off = align256(sizeof(ExchangeHeader));
  for (int b = 0; b < 2; ++b) { g->off_idx[b] = off; off = align256(off + cap_total * 8); }
  for (int b = 0; b < 2; ++b) { g->off_act[b] = off; off = align256(off + cap_total); }
  // Slot area for the tick exchange, sized for ANY split of cap_total records over `world` shards
  // (the split is only known at am_gather_set_layout, after the block has been exported):
  // per parity, the groups / tiles of all shards plus one ragged group per shard and padding.
  {
    const size_t g_tot = (size_t)((cap_total + kGroupRecords - 1) / kGroupRecords) + (size_t)world;
    const size_t t_tot = g_tot * kGroupTiles;
    const size_t per_parity = g_tot * kGroupWords * 4 + (g_tot + (size_t)world) * 4 + t_tot * 4 + t_tot * kTile * 4 +
                              (size_t)world * 4 * 256;
    g->slots_off = off;
    g->slots_bytes = 2 * align256(per_parity);
    off = align256(off + g->slots_bytes);
  }
  g->block_bytes = off;
  int rc = [&]() -> int {
    AMG_CUDA(g, cudaSetDevice(device));
    int sms = 148;
    AMG_CUDA(g, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    g->n_ctas = sms;  // one CTA per SM: the exchange shares the GPU with the next sweep
    if (const char* e = getenv("AMSWEEP_PUSH_CTAS")) { int v = atoi(e); if (v > 0 && v <= 4096) g->n_ctas = v; }
    if (const char* e = getenv("AMSWEEP_PUSH_TIMEOUT_MS")) { long v = atol(e); if (v >= 0) g->push_timeout_ms = (unsigned long long)v; }
    AMG_CUDA(g, cudaMalloc((void**)&g->block, g->block_bytes));
    AMG_CUDA(g, cudaMemset(g->block, 0, g->block_bytes));
    AMG_CUDA(g, cudaMalloc((void**)&g->out_counts, (kMaxWorld + 1) * 4));
    AMG_CUDA(g, cudaMemset(g->out_counts, 0, (kMaxWorld + 1) * 4));
    AMG_CUDA(g, cudaMalloc((void**)&g->status, 4));
    AMG_CUDA(g, cudaMemset(g->status, 0, 4));
    AMG_CUDA(g, cudaDeviceSynchronize());
    return AM_OK;
  }();
  g->peer[rank] = g->block;
  if (rc != AM_OK) { am_gather_destroy(g); return rc; }
  *out = g;
  return AM_OK;
}

int am_gather_export(am_gather_t* g, void* handle_out) {
  if (!g || !handle_out) return AM_E_INVAL;
  static_assert(sizeof(cudaIpcMemHandle_t) == AM_IPC_HANDLE_BYTES, "IPC handle size");
  AMG_CUDA(g, cudaSetDevice(g->device));
  cudaIpcMemHandle_t h;
  AMG_CUDA(g, cudaIpcGetMemHandle(&h, g->block));
  memcpy(handle_out, &h, sizeof h);
  return AM_OK;
}

int am_gather_connect(am_gather_t* g, const void* handles) {
  if (!g || !handles) return AM_E_INVAL;
  AMG_CUDA(g, cudaSetDevice(g->device));
  for (int r = 0; r < g->world; ++r) {
    if (r == g->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const char*)handles + (size_t)r * AM_IPC_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    AMG_CUDA(g, cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    g->peer[r] = (unsigned char*)p;
    g->opened[r] = true;
  }
  g->connected = true;
  return AM_OK;
}

int am_gather_set_layout(am_gather_t* g, const uint64_t* bases, const uint64_t* sizes) {
  if (!g || !bases || !sizes) return AM_E_INVAL;
  uint64_t total = 0;
  for (int r = 0; r < g->world; ++r) {
    if (sizes[r] == 0) return AM_E_INVAL;
    if (g->idx_bytes == 4 && bases[r] + sizes[r] > (1ull << 32)) return AM_E_RANGE;
    total += sizes[r];
  }
  if (total > g->cap_total) return AM_E_RANGE;
  // the same arithmetic on every rank (bases / sizes are all-gathered): identical offsets everywhere
  size_t off = g->slots_off;
  for (int b = 0; b < 2; ++b)
    for (int r = 0; r < g->world; ++r) {
      const uint32_t ng = groups_of(sizes[r]), nt = tiles_of(sizes[r]);
      g->off_bitmap[b][r] = off;   off = align256(off + (size_t)ng * kGroupWords * 4);
      g->off_prefix[b][r] = off;   off = align256(off + ((size_t)ng + 1) * 4);
      g->off_tile_exc[b][r] = off; off = align256(off + (size_t)nt * 4);
      g->off_exc[b][r] = off;      off = align256(off + (size_t)nt * kTile * 4);
    }
  if (off > g->slots_off + g->slots_bytes) return AM_E_RANGE;
  for (int r = 0; r < g->world; ++r) {
    g->bases[r] = bases[r];
    g->sizes[r] = sizes[r];
    g->ngroups[r] = groups_of(sizes[r]);
    g->ntiles[r] = tiles_of(sizes[r]);
  }
  g->layout = true;
  return AM_OK;
}

int am_gather_exchange(am_gather_t* g, am_sweep_t* sweep, void* d_stats, void* cuda_stream) {
  if (!g || !sweep) return AM_E_INVAL;
  if (!g->layout || (!g->connected && g->world > 1)) return AM_E_INVAL;
  ShardTick t{};
  // validate BEFORE the epoch advances: a rank that bumps its epoch without launching would
  // leave its peers waiting for a done flag that never arrives
  if (!shard_last_tick(sweep, &t)) { g->last_error = "am_gather_exchange: no am_sweep_tick_shard on this handle yet"; return AM_E_INVAL; }
  if (t.device != g->device || t.shard_base != g->bases[g->rank] || t.n_records != g->sizes[g->rank]) {
    g->last_error = "am_gather_exchange: the sweep handle is not this rank's shard (device / base / size differ from set_layout)";
    return AM_E_INVAL;
  }
  AMG_CUDA(g, cudaSetDevice(g->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (shard_order_consumer(sweep, st) != AM_OK) { g->last_error = am_last_error_detail(sweep); return AM_E_DEVICE; }
  g->epoch += 1;
  if (g->epoch == 0) g->epoch = 2;  // 0 is the "never written" value of the flag words
  const int buf = g->epoch & 1;
  PushTickParams p{};
  for (int r = 0; r < g->world; ++r) p.peer[r] = g->peer[r];
  p.bitmap = t.out.bitmap;
  p.group_prefix = t.out.group_prefix;
  p.tile_exc = t.out.tile_exc;
  p.exc_seg = t.out.exc_seg;
  p.off_bitmap = g->off_bitmap[buf][g->rank];
  p.off_prefix = g->off_prefix[buf][g->rank];
  p.off_tile_exc = g->off_tile_exc[buf][g->rank];
  p.off_exc = g->off_exc[buf][g->rank];
  p.n_groups = t.n_groups;
  p.n_tiles = t.n_tiles;
  p.epoch = g->epoch;
  p.rank = g->rank;
  p.world = g->world;
  p.timeout_ns = g->push_timeout_ms * 1000000ull;
  p.status = g->status;
  if (g->profiling) AMG_CUDA(g, cudaEventRecord(g->evp[0], st));
  AM_LAUNCH(gather_push_tick_kernel, g->n_ctas, 256, st, p);
  if (g->profiling) AMG_CUDA(g, cudaEventRecord(g->evp[1], st));
  // every rank's slot in MY block is complete when the push retires: rebuild the global list
  ExpandParams e{};
  CountsParams c{};
  uint32_t ng_max = 1;
  for (int r = 0; r < g->world; ++r) {
    ExpandSrc& s = e.src[r];
    if (r == g->rank) {
      s.bitmap = t.out.bitmap; s.group_prefix = t.out.group_prefix; s.tile_exc = t.out.tile_exc; s.exc_seg = t.out.exc_seg;
    } else {
      s.bitmap = reinterpret_cast<const uint32_t*>(g->block + g->off_bitmap[buf][r]);
      s.group_prefix = reinterpret_cast<const uint32_t*>(g->block + g->off_prefix[buf][r]);
      s.tile_exc = reinterpret_cast<const uint32_t*>(g->block + g->off_tile_exc[buf][r]);
      s.exc_seg = reinterpret_cast<const uint32_t*>(g->block + g->off_exc[buf][r]);
    }
    s.base = g->bases[r];
    s.n_groups = g->ngroups[r];
    s.n_tiles = g->ntiles[r];
    c.group_prefix[r] = s.group_prefix;
    c.n_groups[r] = s.n_groups;
    if (s.n_groups > ng_max) ng_max = s.n_groups;
  }
  e.out_idx = g->block + g->off_idx[buf];
  e.out_act = g->block + g->off_act[buf];
  e.acc = t.acc;
  e.stats_base = g->bases[g->rank];
  e.cap = g->cap_total;
  e.world = g->world;
  e.stats_rank = g->rank;
  e.idx_bytes = g->idx_bytes;
  c.out_counts = g->out_counts;
  c.status = g->status;
  c.cap_total = g->cap_total;
  c.world = g->world;
  AM_LAUNCH(gather_counts_kernel, 1, 32, st, c);
  AMG_CUDA(g, cudaGetLastError());
  int rc = shard_launch_expand(sweep, e, ng_max, (uint32_t)g->world, st);
  if (g->profiling) AMG_CUDA(g, cudaEventRecord(g->evp[2], st));
  if (rc == AM_OK) rc = shard_launch_publish(sweep, t.acc, (am_tick_stats_t*)d_stats, t.n_records, st);
  if (g->profiling) { AMG_CUDA(g, cudaEventRecord(g->evp[3], st)); g->profiled = true; }
  if (rc == AM_OK) rc = shard_mark_consumed(sweep, t.parity, st);
  if (rc != AM_OK) g->last_error = am_last_error_detail(sweep);
  return rc;
}

int am_gather_bind(am_gather_t* g, am_sweep_t* shard, void* sweep_stream, void* exchange_stream) {
  if (!g || !shard) return AM_E_INVAL;
  if (!g->layout) return AM_E_INVAL;
  AMG_CUDA(g, cudaSetDevice(g->device));
  const uint64_t cap = g->sizes[g->rank];
  if (!g->view_host || g->view_cap < cap) {
    if (g->view_host) cudaFreeHost(g->view_host);
    g->view_host = nullptr;
    AMG_CUDA(g, cudaHostAlloc((void**)&g->view_host, cap * 5 + 64 + sizeof(am_tick_stats_t), cudaHostAllocMapped));
    g->view_cap = cap;
  }
  if (!g->view_done) AMG_CUDA(g, cudaEventCreateWithFlags(&g->view_done, cudaEventDisableTiming));
  g->bound = shard;
  g->s_sweep = (cudaStream_t)sweep_stream;
  g->s_exchange = (cudaStream_t)exchange_stream;
  return AM_OK;
}

int am_gather_tick_view(am_gather_t* g, int64_t unix_sec, uint32_t mode, am_tick_view_t* view, am_tick_stats_t* stats) {
  if (!g || !view || !g->bound) return AM_E_INVAL;
  AMG_CUDA(g, cudaSetDevice(g->device));
  unsigned char* h = g->view_host;
  unsigned char* d = nullptr;
  AMG_CUDA(g, cudaHostGetDevicePointer((void**)&d, h, 0));
  const size_t off_act = g->view_cap * 4, off_n = (g->view_cap * 5 + 7) / 8 * 8, off_st = off_n + 8;
  int rc = am_sweep_tick_shard(g->bound, unix_sec, mode, g->s_sweep);
  if (rc != AM_OK) { g->last_error = am_last_error_detail(g->bound); return rc; }
  rc = am_gather_exchange(g, g->bound, d + off_st, g->s_exchange);
  if (rc != AM_OK) return rc;
  ExtractParams x{};
  const int buf = g->epoch & 1;
  x.gidx = g->block + g->off_idx[buf];
  x.gact = g->block + g->off_act[buf];
  x.counts = g->out_counts;
  x.out_idx = (uint32_t*)d;
  x.out_act = d + off_act;
  x.out_n = (unsigned long long*)(d + off_n);
  x.shard_base = g->bases[g->rank];
  x.cap = g->view_cap;
  x.rank = g->rank;
  x.world = g->world;
  x.idx_bytes = g->idx_bytes;
  AM_LAUNCH(gather_extract_own_kernel, 148, 256, g->s_exchange, x);
  AMG_CUDA(g, cudaGetLastError());
  AMG_CUDA(g, cudaEventRecord(g->view_done, g->s_exchange));
  AMG_CUDA(g, cudaEventSynchronize(g->view_done));
  const unsigned long long n = *(const volatile unsigned long long*)(h + off_n);
  if ((n >> 32) == kPeerTimeout) { g->last_error = "am_gather_tick_view: a peer did not arrive in time"; return AM_E_DEVICE; }
  view->idx_local = (const uint32_t*)h;
  view->action = h + off_act;
  view->n = n;
  view->shard_base = g->bases[g->rank];
  if (stats) memcpy(stats, h + off_st, sizeof *stats);
  return AM_OK;
}

int am_gather_push(am_gather_t* g, const void* d_idx_local, const void* d_act_local,
                   const void* d_count_local, uint64_t shard_base, void* cuda_stream) {
  if (!g || !d_idx_local || !d_act_local || !d_count_local) return AM_E_INVAL;
  if (!g->connected && g->world > 1) return AM_E_INVAL;
  // validate BEFORE the epoch advances (see am_gather_exchange)
  if (g->idx_bytes == 4 && shard_base > 0xFFFFFFFFull) return AM_E_RANGE;
  if (g->layout && g->idx_bytes == 4 && shard_base + g->sizes[g->rank] > (1ull << 32)) return AM_E_RANGE;
  AMG_CUDA(g, cudaSetDevice(g->device));
  PushParams p{};
  for (int r = 0; r < g->world; ++r) p.peer[r] = g->peer[r];
  p.idx_local = (const uint32_t*)d_idx_local;
  p.act_local = (const uint8_t*)d_act_local;
  p.count_local = (const uint32_t*)d_count_local;
  p.out_counts = g->out_counts;
  p.shard_base = shard_base;
  p.cap_total = g->cap_total;
  for (int b = 0; b < 2; ++b) { p.off_idx[b] = g->off_idx[b]; p.off_act[b] = g->off_act[b]; }
  g->epoch += 1;
  if (g->epoch == 0) g->epoch = 2;
  p.epoch = g->epoch;
  p.rank = g->rank;
  p.world = g->world;
  p.idx_bytes = g->idx_bytes;
  AM_LAUNCH(gather_push_kernel, g->n_ctas, 256, (cudaStream_t)cuda_stream, p);
  AMG_CUDA(g, cudaGetLastError());
  return AM_OK;
}

int am_gather_set_profiling(am_gather_t* g, int on) {
  if (!g) return AM_E_INVAL;
  AMG_CUDA(g, cudaSetDevice(g->device));
  if (on && !g->evp[0])
    for (int k = 0; k < 4; ++k) AMG_CUDA(g, cudaEventCreate(&g->evp[k]));
  g->profiling = on != 0;
  g->profiled = false;
  return AM_OK;
}
int am_gather_last_profile(am_gather_t* g, double* push_ms, double* rebuild_ms, double* publish_ms) {
  if (!g || !g->profiled) return AM_E_INVAL;
  AMG_CUDA(g, cudaSetDevice(g->device));
  AMG_CUDA(g, cudaEventSynchronize(g->evp[3]));
  float a = 0, b = 0, c = 0;
  AMG_CUDA(g, cudaEventElapsedTime(&a, g->evp[0], g->evp[1]));
  AMG_CUDA(g, cudaEventElapsedTime(&b, g->evp[1], g->evp[2]));
  AMG_CUDA(g, cudaEventElapsedTime(&c, g->evp[2], g->evp[3]));
  if (push_ms) *push_ms = a;
  if (rebuild_ms) *rebuild_ms = b;
  if (publish_ms) *publish_ms = c;
  return AM_OK;
}

void* am_gather_out_idx(am_gather_t* g) { return g ? (void*)(g->block + g->off_idx[g->epoch & 1]) : nullptr; }
void* am_gather_out_act(am_gather_t* g) { return g ? g->block + g->off_act[g->epoch & 1] : nullptr; }
void* am_gather_out_counts(am_gather_t* g) { return g ? g->out_counts : nullptr; }
const char* am_gather_last_error(const am_gather_t* g) { return g ? g->last_error.c_str() : ""; }

void am_gather_destroy(am_gather_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < g->world; ++r)
    if (g->opened[r] && g->peer[r]) cudaIpcCloseMemHandle(g->peer[r]);
  if (g->block) cudaFree(g->block);
  if (g->out_counts) cudaFree(g->out_counts);
  if (g->status) cudaFree(g->status);
  for (int k = 0; k < 4; ++k) if (g->evp[k]) cudaEventDestroy(g->evp[k]);
  if (g->view_done) cudaEventDestroy(g->view_done);
  if (g->view_host) cudaFreeHost(g->view_host);
  delete g;
}

}  // extern "C"
