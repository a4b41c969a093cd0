// gather.cu — concatenation of the per-GPU due lists over NVLink peer memory.
//
// The record array shards by contiguous global index range, so
// the global ascending due list is the rank-ordered concatenation of the local
// lists.  NCCL has no allgatherv; the portable path (gather.py) pads to the
// largest count and needs a host round-trip to size the exchange.  Here ONE
// kernel per tick does the whole exchange through CUDA-IPC-mapped peer memory:
//
//   1. every rank stores {epoch, my count} into slot[rank] of every peer's
//      exchange block (system-scope release);
//   2. every CTA waits until all `world` counts of this epoch have arrived in
//      its own block, giving offset = sum(count[r], r < rank) and the total;
//   3. the CTAs stream the local (u32 local index, u8 action) list and write
//      (u32|u64 global index, u8 action) at `offset` into EVERY peer's output
//      buffer, as destination-aligned 16 B + 4 B vector stores — NVSwitch gives
//      each peer full bandwidth;
//   4. the last CTA fences (system scope), raises done[rank] on every peer and
//      waits for all peers' done flags: when the kernel retires, this rank's
//      output buffer holds the complete global list.
//
// Output buffers are double-buffered by epoch parity: a rank can be at most one
// epoch ahead of a peer (step 4), and each rank's consumers are stream-ordered
// before its next push, so epoch e+2 never overwrites data still being read.
//
// The reference has no counterpart (single Go process, no collectives); the
// consumer of the list is createSubmitWorkflow, hcc.go:502.
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

struct am_gather {
  int device = 0, rank = 0, world = 1;
  uint64_t cap_total = 0;
  int idx_bytes = 8;
  int n_ctas = 296;
  size_t block_bytes = 0;
  size_t off_idx[2] = {0, 0}, off_act[2] = {0, 0};
  unsigned char* block = nullptr;                 // my exchange block (cudaMalloc, IPC-exported)
  unsigned char* peer[kMaxWorld] = {};            // mapped peers (own = block)
  bool opened[kMaxWorld] = {};
  uint32_t* out_counts = nullptr;
  uint32_t epoch = 0;
  bool connected = false;
  // compressed wire format (enabled by am_gather_set_layout)
  bool compressed = false;
  size_t off_gc[2] = {0, 0}, off_o16[2] = {0, 0};
  uint32_t ngroups_max = 0;
  uint64_t bases[kMaxWorld] = {}, sizes[kMaxWorld] = {};
  uint32_t ngroups[kMaxWorld] = {};
  void* final_idx[2] = {nullptr, nullptr};  // expanded global indices, by epoch parity
  // bitmap wire format (am_gather_set_wire(AM_WIRE_BITMAP) after set_layout; experimental)
  int wire = AM_WIRE_PLAIN;
  unsigned long long push_timeout_ms = 5000;  // bitmap format only (AMSWEEP_PUSH_TIMEOUT_MS, 0 = none)
  size_t off_bm[2] = {0, 0};
  uint64_t bm_word0[kMaxWorld] = {};
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
  auto align = [](size_t v) { return (v + 255) / 256 * 256; };
  size_t off = align(sizeof(ExchangeHeader));
  for (int b = 0; b < 2; ++b) { g->off_idx[b] = off; off = align(off + cap_total * 8); }
  for (int b = 0; b < 2; ++b) { g->off_act[b] = off; off = align(off + cap_total); }
  // compressed wire format: per-group counts [world][ngroups_max] and u16 offsets [cap_total]
  g->ngroups_max = (uint32_t)((cap_total + kGroupRecords - 1) / kGroupRecords);
  for (int b = 0; b < 2; ++b) { g->off_gc[b] = off; off = align(off + (size_t)world * g->ngroups_max * 4); }
  for (int b = 0; b < 2; ++b) { g->off_o16[b] = off; off = align(off + cap_total * 2); }
  // bitmap wire format: one 1 KB bitmap per 8192-record group of every shard (at most
  // ngroups_max + world groups in total, whatever the split)
  for (int b = 0; b < 2; ++b) {
    g->off_bm[b] = off;
    off = align(off + ((size_t)g->ngroups_max + (size_t)world) * kGroupWords * 4);
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
  AMG_CUDA(g, cudaSetDevice(g->device));
  for (int r = 0; r < g->world; ++r) {
    const uint64_t ng = (sizes[r] + kGroupRecords - 1) / kGroupRecords;
    if (ng > g->ngroups_max) return AM_E_RANGE;
    if (g->idx_bytes == 4 && bases[r] + sizes[r] > (1ull << 32)) return AM_E_RANGE;
    g->bases[r] = bases[r];
    g->sizes[r] = sizes[r];
    g->ngroups[r] = (uint32_t)ng;
  }
  uint64_t groups_before = 0;
  for (int r = 0; r < g->world; ++r) {
    g->bm_word0[r] = groups_before * kGroupWords;
    groups_before += g->ngroups[r];
  }
  if (groups_before > (uint64_t)g->ngroups_max + (uint64_t)g->world) return AM_E_RANGE;
  for (int b = 0; b < 2; ++b)
    if (!g->final_idx[b]) AMG_CUDA(g, cudaMalloc(&g->final_idx[b], g->cap_total * (size_t)g->idx_bytes));
  g->compressed = true;
  g->wire = AM_WIRE_C3;
  return AM_OK;
}

int am_gather_set_wire(am_gather_t* g, int wire) {
  if (!g) return AM_E_INVAL;
  if (wire != AM_WIRE_C3 && wire != AM_WIRE_BITMAP) return AM_E_INVAL;  // plain = never call set_layout
  if (!g->compressed) return AM_E_INVAL;                              // needs the shard layout
  g->wire = wire;
  return AM_OK;
}

int am_gather_push(am_gather_t* g, const void* d_idx_local, const void* d_act_local,
                   const void* d_count_local, uint64_t shard_base, void* cuda_stream) {
  if (!g || !d_idx_local || !d_act_local || !d_count_local) return AM_E_INVAL;
  if (!g->connected && g->world > 1) return AM_E_INVAL;
  // validate BEFORE the epoch advances: a rank that bumps its epoch without launching
  // would leave its peers spinning on counts that never arrive
  if (g->compressed && shard_base != g->bases[g->rank]) return AM_E_INVAL;
  if (!g->compressed && g->idx_bytes == 4 && shard_base > 0xFFFFFFFFull) return AM_E_RANGE;
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
  if (g->epoch == 0) g->epoch = 2;  // keep parity continuity irrelevant: 0 is the "never written" value
  if (g->compressed && g->wire == AM_WIRE_BITMAP) {
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int buf = g->epoch & 1;
    // default actions: stream-ordered before this rank publishes its count, hence before any
    // peer stores a non-default action of this epoch into the buffer
    AMG_CUDA(g, cudaMemsetAsync(g->block + g->off_act[buf], (int)AM_ACT_SUBMIT_HC, g->cap_total, st));
    PushBmParams b{};
    for (int r = 0; r < g->world; ++r) b.peer[r] = g->peer[r];
    b.idx_local = (const uint32_t*)d_idx_local;
    b.act_local = (const uint8_t*)d_act_local;
    b.count_local = (const uint32_t*)d_count_local;
    b.out_counts = g->out_counts;
    b.cap_total = g->cap_total;
    for (int k = 0; k < 2; ++k) { b.off_act[k] = g->off_act[k]; b.off_gc[k] = g->off_gc[k]; b.off_bm[k] = g->off_bm[k]; }
    b.bm_word0 = g->bm_word0[g->rank];
    b.timeout_ns = g->push_timeout_ms * 1000000ull;
    b.epoch = g->epoch;
    b.ngroups_mine = g->ngroups[g->rank];
    b.ngroups_max = g->ngroups_max;
    b.rank = g->rank;
    b.world = g->world;
    AM_LAUNCH(gather_push_bm_kernel, g->n_ctas, 256, st, b);
    ExpandBmParams x{};
    x.bm = reinterpret_cast<const uint32_t*>(g->block + g->off_bm[buf]);
    x.gc = reinterpret_cast<const uint32_t*>(g->block + g->off_gc[buf]);
    x.counts = g->out_counts;
    x.final_idx = g->final_idx[buf];
    x.cap_total = g->cap_total;
    x.ngroups_max = g->ngroups_max;
    x.world = g->world;
    x.idx_bytes = g->idx_bytes;
    uint32_t ng_used = 1;
    for (int r = 0; r < g->world; ++r) {
      x.ngroups[r] = g->ngroups[r];
      x.bases[r] = g->bases[r];
      x.bm_word0[r] = g->bm_word0[r];
      if (g->ngroups[r] > ng_used) ng_used = g->ngroups[r];
    }
    AM_LAUNCH(gather_expand_bitmap_kernel, dim3(ng_used, g->world), 256, st, x);
    AMG_CUDA(g, cudaGetLastError());
    return AM_OK;
  }
  if (g->compressed) {
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int buf = g->epoch & 1;
    PushC3Params c{};
    for (int r = 0; r < g->world; ++r) c.peer[r] = g->peer[r];
    c.idx_local = (const uint32_t*)d_idx_local;
    c.act_local = (const uint8_t*)d_act_local;
    c.count_local = (const uint32_t*)d_count_local;
    c.out_counts = g->out_counts;
    c.cap_total = g->cap_total;
    for (int b = 0; b < 2; ++b) { c.off_act[b] = g->off_act[b]; c.off_gc[b] = g->off_gc[b]; c.off_o16[b] = g->off_o16[b]; }
    c.epoch = g->epoch;
    c.ngroups_mine = g->ngroups[g->rank];
    c.ngroups_max = g->ngroups_max;
    c.rank = g->rank;
    c.world = g->world;
    AM_LAUNCH(gather_push_c3_kernel, g->n_ctas, 256, st, c);
    DecodeParams dp{};
    dp.o16 = reinterpret_cast<const uint16_t*>(g->block + g->off_o16[buf]);
    dp.gc = reinterpret_cast<const uint32_t*>(g->block + g->off_gc[buf]);
    dp.counts = g->out_counts;
    dp.final_idx = g->final_idx[buf];
    dp.cap_total = g->cap_total;
    dp.ngroups_max = g->ngroups_max;
    dp.world = g->world;
    dp.idx_bytes = g->idx_bytes;
    uint32_t ng_used = 1;
    for (int r = 0; r < g->world; ++r) {
      dp.ngroups[r] = g->ngroups[r];
      dp.bases[r] = g->bases[r];
      if (g->ngroups[r] > ng_used) ng_used = g->ngroups[r];
    }
    AM_LAUNCH(gather_decode_kernel, dim3(ng_used, g->world), 256, st, dp);
    AMG_CUDA(g, cudaGetLastError());
    return AM_OK;
  }
  p.epoch = g->epoch;
  p.rank = g->rank;
  p.world = g->world;
  p.idx_bytes = g->idx_bytes;
  AM_LAUNCH(gather_push_kernel, g->n_ctas, 256, (cudaStream_t)cuda_stream, p);
  AMG_CUDA(g, cudaGetLastError());
  return AM_OK;
}

void* am_gather_out_idx(am_gather_t* g) {
  if (!g) return nullptr;
  return g->compressed ? g->final_idx[g->epoch & 1] : (void*)(g->block + g->off_idx[g->epoch & 1]);
}
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
  for (int b = 0; b < 2; ++b) if (g->final_idx[b]) cudaFree(g->final_idx[b]);
  delete g;
}

}  // extern "C"
