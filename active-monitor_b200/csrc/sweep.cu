// sweep.cu — the am_sweep handle and the C-ABI entry points of libamsweep.
//
// Owns one shard of the HealthCheck record array in HBM (16 SoA columns,
// SURVEY.md Appendix B.1) on one CUDA device, the staging area for
// upsert / remove / post_result calls coming from the controller's goroutines
// (hcc.go:170-188 Reconcile workers, hcc.go:635/:662/:821/:836 watch loops),
// and the launch of the sweep kernel (sweep_kernels.cuh) that replaces the
// per-CR schedule ladder and remedy state machine for every record at once.
//
// There is deliberately no CPU path in this file: without a CUDA device
// am_sweep_create fails with AM_E_DEVICE.
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "sweep_kernels.cuh"

using namespace amsweep;

namespace {
// Reason of the last failed am_sweep_create, process-wide: a cgo caller may be moved to
// another OS thread between the failing call and am_last_error_detail(NULL).
std::mutex g_create_mu;
std::string g_create_error;
void set_create_error(const std::string& s) {
  std::lock_guard<std::mutex> lk(g_create_mu);
  g_create_error = s;
}

struct PinnedBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 2 + 4096;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 2 + 4096;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};
}  // namespace

struct am_sweep {
  int device = 0;
  uint64_t capacity = 0, cap_padded = 0, shard_base = 0;
  uint64_t n_records = 0;  // high-water mark
  DevCols cols{};
  void* col_ptr[16] = {};
  size_t col_elem[16] = {};
  uint32_t* seg_idx = nullptr;       // per-tile segments written by the sweep kernel
  uint8_t* seg_act = nullptr;
  uint32_t* tile_count = nullptr;    // [tiles]
  uint32_t* group_count[2] = {nullptr, nullptr};  // [groups], parity = tick number & 1
  unsigned long long* acc = nullptr;
  uint32_t parity = 0;
  uint32_t* due_idx[2] = {nullptr, nullptr};
  uint8_t* due_action[2] = {nullptr, nullptr};
  am_tick_stats_t* h_stats = nullptr;  // pinned + mapped
  am_tick_stats_t* d_stats_mapped = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  cudaEvent_t evp[3] = {nullptr, nullptr, nullptr};  // profiling: before sweep, between, after compact
  bool profiling = false, profiled = false;
  std::mutex mu;  // guards the staged vectors
  std::atomic_flag ticking = ATOMIC_FLAG_INIT;
  // staged controller events in arrival order, written straight into pinned host
  // memory (the H2D copy at the next tick reads them in place)
  PinnedBuf st_ops, st_recs;
  size_t n_ops = 0, n_recs = 0;
  uint32_t n_state_ops = 0, n_result_ops = 0;
  uint32_t* marks = nullptr;  // [2 * cap_padded] per-slot {latest state op, latest result} of this tick
  PinnedBuf pin_in, pin_out;
  DevBuf dev_in;
  std::string last_error;
  uint64_t launches = 0;
  double last_ms = -1.0;
  uint64_t seed = 0;
};

namespace {

#define AM_CUDA(h, expr)                                                              \
  do {                                                                                \
    cudaError_t _e = (expr);                                                          \
    if (_e != cudaSuccess) {                                                          \
      char _b[512];                                                                   \
      snprintf(_b, sizeof _b, "%s -> %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      if (h) (h)->last_error = _b; else set_create_error(_b);                         \
      return _e == cudaErrorMemoryAllocation ? AM_E_NOMEM : AM_E_DEVICE;              \
    }                                                                                 \
  } while (0)

constexpr size_t kColElem[16] = {8, 8, 8, 8, 8, 4, 4, 8, 4, 4, 4, 4, 4, 4, 4, 8};

void** col_slot(DevCols& c, int k) {
  void** slots[16] = {(void**)&c.minute,         (void**)&c.hour,          (void**)&c.dom,
                      (void**)&c.month,          (void**)&c.dow,           (void**)&c.ras,
                      (void**)&c.flags,          (void**)&c.finished_at,   (void**)&c.runs_limit,
                      (void**)&c.reset_interval, (void**)&c.success,       (void**)&c.failed,
                      (void**)&c.remedy_success, (void**)&c.remedy_failed, (void**)&c.remedy_total,
                      (void**)&c.remedy_finished_at};
  return slots[k];
}
void* const* cols_member(const am_record_cols_t* c, int k) {
  void* const* slots[16] = {(void* const*)&c->minute,         (void* const*)&c->hour,
                            (void* const*)&c->dom,            (void* const*)&c->month,
                            (void* const*)&c->dow,            (void* const*)&c->ras,
                            (void* const*)&c->flags,          (void* const*)&c->finished_at,
                            (void* const*)&c->runs_limit,     (void* const*)&c->reset_interval,
                            (void* const*)&c->success,        (void* const*)&c->failed,
                            (void* const*)&c->remedy_success, (void* const*)&c->remedy_failed,
                            (void* const*)&c->remedy_total,   (void* const*)&c->remedy_finished_at};
  return slots[k];
}

struct TickGuard {
  am_sweep* h;
  bool ok;
  explicit TickGuard(am_sweep* hh) : h(hh), ok(!hh->ticking.test_and_set(std::memory_order_acquire)) {}
  ~TickGuard() { if (ok) h->ticking.clear(std::memory_order_release); }
};

// Grow a pinned staging array (called under h->mu); keeps the contents.
cudaError_t grow_pinned(PinnedBuf& b, size_t used_bytes, size_t want_bytes) {
  if (want_bytes <= b.cap) return cudaSuccess;
  void* np = nullptr;
  size_t ncap = want_bytes * 2 + 65536;
  cudaError_t e = cudaHostAlloc(&np, ncap, cudaHostAllocDefault);
  if (e != cudaSuccess) return e;
  if (used_bytes) memcpy(np, b.p, used_bytes);
  if (b.p) cudaFreeHost(b.p);
  b.p = np;
  b.cap = ncap;
  return cudaSuccess;
}

// Apply staged upserts / removes / results (called with the tick guard held).
int drain_staged(am_sweep* h) {
  // Hold the staging lock for the H2D enqueue only: the copies read the pinned arrays in
  // place, and callers may not append until the copy has been consumed.
  std::unique_lock<std::mutex> lk(h->mu);
  const size_t n = h->n_ops, nrec = h->n_recs;
  if (n == 0) return AM_OK;
  const uint32_t n_state = h->n_state_ops, n_result = h->n_result_ops;
  const StagedOp* ops = (const StagedOp*)h->st_ops.p;
  if (n_state)
    for (size_t k = 0; k < n; ++k)  // every upserted slot extends the swept range (high-water mark)
      if ((ops[k].arg & kOpKindMask) == kOpUpsert && (uint64_t)ops[k].idx + 1 > h->n_records)
        h->n_records = (uint64_t)ops[k].idx + 1;
  const size_t o_rec = (n * sizeof(StagedOp) + 255) / 256 * 256, total = o_rec + nrec * sizeof(am_record_t);
  AM_CUDA(h, h->dev_in.reserve(total));
  char* dp = (char*)h->dev_in.p;
  AM_CUDA(h, cudaMemcpyAsync(dp, h->st_ops.p, n * sizeof(StagedOp), cudaMemcpyHostToDevice, h->stream));
  if (nrec) AM_CUDA(h, cudaMemcpyAsync(dp + o_rec, h->st_recs.p, nrec * sizeof(am_record_t), cudaMemcpyHostToDevice, h->stream));
  const StagedOp* d_ops = (const StagedOp*)dp;
  const am_record_t* d_recs = (const am_record_t*)(dp + o_rec);
  const unsigned B = 256, G = (unsigned)((n + B - 1) / B);
  AM_LAUNCH(mark_ops_kernel, G, B, h->stream, h->marks, d_ops, (uint32_t)n);
  h->launches++;
  if (n_state) {
    AM_LAUNCH(apply_state_ops_kernel, G, B, h->stream, h->cols, h->marks, d_ops, d_recs, (uint32_t)n);
    h->launches++;
  }
  if (n_result) {
    AM_LAUNCH(apply_result_ops_kernel, G, B, h->stream, h->cols.flags, h->marks, d_ops, (uint32_t)n);
    h->launches++;
  }
  AM_LAUNCH(clear_marks_kernel, G, B, h->stream, h->marks, d_ops, (uint32_t)n);
  h->launches++;
  AM_CUDA(h, cudaGetLastError());
  // the pinned arrays are refilled by the next calls: wait until the copies were consumed
  AM_CUDA(h, cudaStreamSynchronize(h->stream));
  h->n_ops = h->n_recs = 0;
  h->n_state_ops = h->n_result_ops = 0;
  return AM_OK;
}

int launch_sweep(am_sweep* h, int64_t T, uint32_t mode, uint32_t* d_idx, uint8_t* d_act, uint64_t cap,
                 am_tick_stats_t* out_stats, uint32_t* out_count, cudaStream_t s) {
  if (h->n_records == 0) {
    // nothing to sweep: publish zeros without a launch
    if (out_stats) AM_CUDA(h, cudaMemsetAsync(out_stats, 0, sizeof(am_tick_stats_t), s));
    if (out_count) AM_CUDA(h, cudaMemsetAsync(out_count, 0, 4, s));
    return AM_OK;
  }
  SweepParams p{};
  p.c = h->cols;
  p.n_records = h->n_records;
  p.shard_base = h->shard_base;
  p.seed = h->seed;
  p.T = T;
  p.words = tick_words_from_unix(T);
  p.n_tiles = (uint32_t)((h->n_records + kTile - 1) / kTile);
  p.mode = mode;
  p.seg_idx = h->seg_idx;
  p.seg_act = h->seg_act;
  p.tile_count = h->tile_count;
  p.group_count = h->group_count[h->parity];
  p.acc = h->acc;
  if (h->profiling) AM_CUDA(h, cudaEventRecord(h->evp[0], s));
  // off the minute no 5-field schedule can fire: the mask columns are not read
  int64_t sec_of_min = T % 60;
  if (sec_of_min < 0) sec_of_min += 60;
  const bool masks = sec_of_min == 0 || (mode & AM_SWEEP_FULL_SCAN);
  const bool closed = (mode & AM_SWEEP_CLOSED_LOOP) != 0;
  if (closed && masks) AM_LAUNCH(AM_SWEEP_KERNEL(true, true), p.n_tiles, kBlock, s, p);
  else if (closed) AM_LAUNCH(AM_SWEEP_KERNEL(true, false), p.n_tiles, kBlock, s, p);
  else if (masks) AM_LAUNCH(AM_SWEEP_KERNEL(false, true), p.n_tiles, kBlock, s, p);
  else AM_LAUNCH(AM_SWEEP_KERNEL(false, false), p.n_tiles, kBlock, s, p);
  if (h->profiling) AM_CUDA(h, cudaEventRecord(h->evp[1], s));
  CompactParams c{};
  c.seg_idx = h->seg_idx;
  c.seg_act = h->seg_act;
  c.tile_count = h->tile_count;
  c.group_count = h->group_count[h->parity];
  c.group_count_next = h->group_count[h->parity ^ 1];
  c.acc = h->acc;
  c.out_idx = d_idx;
  c.out_act = d_act;
  c.shard_base = h->shard_base;
  c.n_tiles = p.n_tiles;
  c.n_groups = (p.n_tiles + kGroupTiles - 1) / kGroupTiles;
  c.cap = (uint32_t)(cap > 0xFFFFFFFFull ? 0xFFFFFFFFull : cap);
  AM_LAUNCH(compact_kernel, c.n_groups, 256, s, c);
  AM_LAUNCH(publish_kernel, 1, 32, s, h->acc, out_stats, out_count, h->n_records);
  if (h->profiling) { AM_CUDA(h, cudaEventRecord(h->evp[2], s)); h->profiled = true; }
  h->parity ^= 1;
  h->launches += 3;
  AM_CUDA(h, cudaGetLastError());
  return AM_OK;
}

}  // namespace

extern "C" {

const char* am_last_error_detail(const am_sweep_t* h) {
  if (h) return h->last_error.c_str();
  static thread_local std::string copy;  // stable storage for the returned pointer
  std::lock_guard<std::mutex> lk(g_create_mu);
  copy = g_create_error;
  return copy.c_str();
}

int am_sweep_create(am_sweep_t** out, int device_id, uint64_t capacity, uint64_t shard_base) {
  if (!out || capacity == 0 || capacity > 0xFFFFF000ull) return AM_E_INVAL;
  *out = nullptr;
  am_sweep* none = nullptr;
  int ndev = 0;
  AM_CUDA(none, cudaGetDeviceCount(&ndev));
  if (device_id < 0 || device_id >= ndev) {
    set_create_error("no such CUDA device");
    return AM_E_DEVICE;
  }
  AM_CUDA(none, cudaSetDevice(device_id));
  am_sweep* h = new (std::nothrow) am_sweep();
  if (!h) return AM_E_NOMEM;
  h->device = device_id;
  h->capacity = capacity;
  h->cap_padded = (capacity + kTile - 1) / kTile * kTile;
  h->shard_base = shard_base;
  int rc = [&]() -> int {
    AM_CUDA(h, cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));
    AM_CUDA(h, cudaEventCreate(&h->ev0));
    AM_CUDA(h, cudaEventCreate(&h->ev1));
    for (int k = 0; k < 3; ++k) AM_CUDA(h, cudaEventCreate(&h->evp[k]));
    for (int k = 0; k < 16; ++k) {
      void* p = nullptr;
      AM_CUDA(h, cudaMalloc(&p, h->cap_padded * kColElem[k]));
      AM_CUDA(h, cudaMemsetAsync(p, 0, h->cap_padded * kColElem[k], h->stream));
      *col_slot(h->cols, k) = p;
      h->col_ptr[k] = p;
      h->col_elem[k] = kColElem[k];
    }
    AM_LAUNCH(fill_u32_kernel, 1184, 256, h->stream, h->cols.flags, AM_F_TOMBSTONE, h->cap_padded);
    h->launches++;
    const size_t ntiles = h->cap_padded / kTile;
    const size_t ngroups = (ntiles + kGroupTiles - 1) / kGroupTiles;
    AM_CUDA(h, cudaMalloc((void**)&h->seg_idx, h->cap_padded * 4));
    AM_CUDA(h, cudaMalloc((void**)&h->seg_act, h->cap_padded));
    AM_CUDA(h, cudaMalloc((void**)&h->tile_count, ntiles * 4));
    AM_CUDA(h, cudaMemsetAsync(h->tile_count, 0, ntiles * 4, h->stream));
    for (int b = 0; b < 2; ++b) {
      AM_CUDA(h, cudaMalloc((void**)&h->group_count[b], ngroups * 4));
      AM_CUDA(h, cudaMemsetAsync(h->group_count[b], 0, ngroups * 4, h->stream));
    }
    AM_CUDA(h, cudaMalloc((void**)&h->acc, kNumAcc * 8));
    AM_CUDA(h, cudaMemsetAsync(h->acc, 0, kNumAcc * 8, h->stream));
    AM_CUDA(h, cudaMalloc((void**)&h->marks, h->cap_padded * 8));
    AM_CUDA(h, cudaMemsetAsync(h->marks, 0, h->cap_padded * 8, h->stream));
    for (int b = 0; b < 2; ++b) {
      AM_CUDA(h, cudaMalloc((void**)&h->due_idx[b], h->cap_padded * 4));
      AM_CUDA(h, cudaMalloc((void**)&h->due_action[b], h->cap_padded));
    }
    AM_CUDA(h, cudaHostAlloc((void**)&h->h_stats, sizeof(am_tick_stats_t), cudaHostAllocMapped));
    memset(h->h_stats, 0, sizeof(am_tick_stats_t));
    AM_CUDA(h, cudaHostGetDevicePointer((void**)&h->d_stats_mapped, h->h_stats, 0));
    AM_CUDA(h, cudaGetLastError());
    AM_CUDA(h, cudaStreamSynchronize(h->stream));
    return AM_OK;
  }();
  if (rc != AM_OK) {
    set_create_error(h->last_error);
    am_sweep_destroy(h);
    return rc;
  }
  *out = h;
  return AM_OK;
}

void am_sweep_destroy(am_sweep_t* h) {
  if (!h) return;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  for (int k = 0; k < 16; ++k) if (h->col_ptr[k]) cudaFree(h->col_ptr[k]);
  if (h->seg_idx) cudaFree(h->seg_idx);
  if (h->seg_act) cudaFree(h->seg_act);
  if (h->tile_count) cudaFree(h->tile_count);
  for (int b = 0; b < 2; ++b) if (h->group_count[b]) cudaFree(h->group_count[b]);
  if (h->acc) cudaFree(h->acc);
  if (h->marks) cudaFree(h->marks);
  for (int b = 0; b < 2; ++b) {
    if (h->due_idx[b]) cudaFree(h->due_idx[b]);
    if (h->due_action[b]) cudaFree(h->due_action[b]);
  }
  if (h->h_stats) cudaFreeHost(h->h_stats);
  h->pin_in.release(); h->pin_out.release(); h->dev_in.release();
  h->st_ops.release(); h->st_recs.release();
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  for (int k = 0; k < 3; ++k) if (h->evp[k]) cudaEventDestroy(h->evp[k]);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
}

uint64_t am_sweep_size(const am_sweep_t* h) { return h ? h->n_records : 0; }
uint64_t am_sweep_capacity(const am_sweep_t* h) { return h ? h->capacity : 0; }
int am_sweep_device(const am_sweep_t* h) { return h ? h->device : -1; }
double am_sweep_last_kernel_ms(const am_sweep_t* h) { return h ? h->last_ms : -1.0; }
uint64_t am_sweep_launch_count(const am_sweep_t* h) { return h ? h->launches : 0; }
void* am_sweep_column_ptr(am_sweep_t* h, int column) {
  return (h && column >= 0 && column < 16) ? h->col_ptr[column] : nullptr;
}
int am_sweep_set_profiling(am_sweep_t* h, int on) {
  if (!h) return AM_E_INVAL;
  h->profiling = on != 0;
  h->profiled = false;
  return AM_OK;
}
int am_sweep_last_profile(am_sweep_t* h, double* sweep_ms, double* compact_ms) {
  if (!h || !h->profiled) return AM_E_INVAL;
  AM_CUDA(h, cudaSetDevice(h->device));
  AM_CUDA(h, cudaEventSynchronize(h->evp[2]));
  float a = 0, b = 0;
  AM_CUDA(h, cudaEventElapsedTime(&a, h->evp[0], h->evp[1]));
  AM_CUDA(h, cudaEventElapsedTime(&b, h->evp[1], h->evp[2]));
  if (sweep_ms) *sweep_ms = a;
  if (compact_ms) *compact_ms = b;
  return AM_OK;
}
void* am_sweep_stream(am_sweep_t* h) { return h ? (void*)h->stream : nullptr; }
int am_sweep_set_seed(am_sweep_t* h, uint64_t seed) {
  if (!h) return AM_E_INVAL;
  h->seed = seed;
  return AM_OK;
}

int am_sweep_load_range(am_sweep_t* h, uint64_t first, uint64_t n, const am_record_cols_t* cols) {
  if (!h || !cols || first + n > h->capacity || first + n < first) return AM_E_INVAL;
  if (n == 0) return AM_OK;
  TickGuard g(h);
  if (!g.ok) return AM_E_BUSY;
  AM_CUDA(h, cudaSetDevice(h->device));
  for (int k = 0; k < 16; ++k) {
    char* dst = (char*)h->col_ptr[k] + first * kColElem[k];
    const void* src = *cols_member(cols, k);
    if (src) AM_CUDA(h, cudaMemcpyAsync(dst, src, n * kColElem[k], cudaMemcpyHostToDevice, h->stream));
    else AM_CUDA(h, cudaMemsetAsync(dst, 0, n * kColElem[k], h->stream));
  }
  AM_CUDA(h, cudaStreamSynchronize(h->stream));
  if (first + n > h->n_records) h->n_records = first + n;
  return AM_OK;
}

int am_sweep_upsert(am_sweep_t* h, uint64_t n, const uint64_t* idx, const am_record_t* recs) {
  if (!h || (n && (!idx || !recs))) return AM_E_INVAL;
  for (uint64_t k = 0; k < n; ++k)
    if (idx[k] >= h->capacity) return AM_E_RANGE;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->n_ops + n > 0x3FFFFFF0ull || h->n_recs + n > 0x3FFFFFF0ull) return AM_E_NOSPACE;
  AM_CUDA(h, cudaSetDevice(h->device));
  AM_CUDA(h, grow_pinned(h->st_ops, h->n_ops * sizeof(StagedOp), (h->n_ops + n) * sizeof(StagedOp)));
  AM_CUDA(h, grow_pinned(h->st_recs, h->n_recs * sizeof(am_record_t), (h->n_recs + n) * sizeof(am_record_t)));
  StagedOp* ops = (StagedOp*)h->st_ops.p + h->n_ops;
  am_record_t* dst = (am_record_t*)h->st_recs.p + h->n_recs;
  for (uint64_t k = 0; k < n; ++k) {
    ops[k] = StagedOp{(uint32_t)idx[k], kOpUpsert | (uint32_t)(h->n_recs + k)};
    dst[k] = recs[k];
    dst[k].flags &= ~AM_F_TOMBSTONE;
    dst[k].reserved = 0;
  }
  h->n_ops += n; h->n_recs += n;
  h->n_state_ops += (uint32_t)n;
  return AM_OK;
}

int am_sweep_remove(am_sweep_t* h, uint64_t n, const uint64_t* idx) {
  if (!h || (n && !idx)) return AM_E_INVAL;
  for (uint64_t k = 0; k < n; ++k)
    if (idx[k] >= h->capacity) return AM_E_RANGE;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->n_ops + n > 0x3FFFFFF0ull) return AM_E_NOSPACE;
  AM_CUDA(h, cudaSetDevice(h->device));
  AM_CUDA(h, grow_pinned(h->st_ops, h->n_ops * sizeof(StagedOp), (h->n_ops + n) * sizeof(StagedOp)));
  StagedOp* ops = (StagedOp*)h->st_ops.p + h->n_ops;
  for (uint64_t k = 0; k < n; ++k) ops[k] = StagedOp{(uint32_t)idx[k], kOpRemove};
  h->n_ops += n;
  h->n_state_ops += (uint32_t)n;
  return AM_OK;
}

int am_sweep_post_result(am_sweep_t* h, uint64_t n, const uint64_t* idx, const uint8_t* phase,
                         const uint8_t* remedy_phase) {
  if (!h || (n && (!idx || !phase))) return AM_E_INVAL;
  std::lock_guard<std::mutex> lk(h->mu);
  if (h->n_ops + n > 0x3FFFFFF0ull) return AM_E_NOSPACE;
  AM_CUDA(h, cudaSetDevice(h->device));
  AM_CUDA(h, grow_pinned(h->st_ops, h->n_ops * sizeof(StagedOp), (h->n_ops + n) * sizeof(StagedOp)));
  StagedOp* ops = (StagedOp*)h->st_ops.p + h->n_ops;
  // one pass: validate and stage; nothing is committed (n_ops unchanged) on a bad entry.
  // phase -> flag bits: {none, Succeeded, Failed} -> {0, PENDING_OK, PENDING_FAIL}
  static const uint32_t kPhaseBits[4] = {0u, AM_F_PENDING_OK, AM_F_PENDING_FAIL, 0u};
  static const uint32_t kRemedyBits[4] = {0u, AM_F_REMEDY_PENDING | AM_F_REMEDY_OUTCOME_OK, AM_F_REMEDY_PENDING, 0u};
  const uint64_t cap = h->capacity;
  uint64_t bad_range = 0, bad_phase = 0;
  if (remedy_phase) {
    for (uint64_t k = 0; k < n; ++k) {
      bad_range |= (uint64_t)(idx[k] >= cap);
      bad_phase |= (uint64_t)(phase[k] > AM_PHASE_FAILED) | (uint64_t)(remedy_phase[k] > AM_PHASE_FAILED);
      ops[k] = StagedOp{(uint32_t)idx[k], kOpResult | kPhaseBits[phase[k] & 3] | kRemedyBits[remedy_phase[k] & 3]};
    }
  } else {
    for (uint64_t k = 0; k < n; ++k) {
      bad_range |= (uint64_t)(idx[k] >= cap);
      bad_phase |= (uint64_t)(phase[k] > AM_PHASE_FAILED);
      ops[k] = StagedOp{(uint32_t)idx[k], kOpResult | kPhaseBits[phase[k] & 3]};
    }
  }
  if (bad_range) return AM_E_RANGE;
  if (bad_phase) return AM_E_INVAL;
  h->n_ops += n;
  h->n_result_ops += (uint32_t)n;
  return AM_OK;
}

int am_sweep_tick(am_sweep_t* h, int64_t unix_sec, uint32_t mode, uint64_t* due_idx,
                  uint32_t* due_action, uint64_t cap, uint64_t* n_out, am_tick_stats_t* stats) {
  if (!h || (cap && (!due_idx || !due_action))) return AM_E_INVAL;
  if (unix_sec >= (1ll << 55) || unix_sec <= -(1ll << 55)) return AM_E_RANGE;
  TickGuard g(h);
  if (!g.ok) return AM_E_BUSY;
  AM_CUDA(h, cudaSetDevice(h->device));
  int rc = drain_staged(h);
  if (rc != AM_OK) return rc;
  AM_CUDA(h, cudaEventRecord(h->ev0, h->stream));
  rc = launch_sweep(h, unix_sec, mode, h->due_idx[0], h->due_action[0], h->cap_padded,
                    h->d_stats_mapped, nullptr, h->stream);
  if (rc != AM_OK) return rc;
  AM_CUDA(h, cudaEventRecord(h->ev1, h->stream));
  AM_CUDA(h, cudaStreamSynchronize(h->stream));
  float ms = 0;
  AM_CUDA(h, cudaEventElapsedTime(&ms, h->ev0, h->ev1));
  h->last_ms = ms;
  am_tick_stats_t st = *h->h_stats;
  st.n_records = h->n_records;
  if (stats) *stats = st;
  const uint64_t n = st.n_emitted;
  if (n_out) *n_out = n;
  const uint64_t ncopy = n < cap ? n : cap;
  if (ncopy) {
    AM_CUDA(h, h->pin_out.reserve(ncopy * 5));
    uint32_t* hi = (uint32_t*)h->pin_out.p;
    uint8_t* ha = (uint8_t*)h->pin_out.p + ncopy * 4;
    AM_CUDA(h, cudaMemcpyAsync(hi, h->due_idx[0], ncopy * 4, cudaMemcpyDeviceToHost, h->stream));
    AM_CUDA(h, cudaMemcpyAsync(ha, h->due_action[0], ncopy, cudaMemcpyDeviceToHost, h->stream));
    AM_CUDA(h, cudaStreamSynchronize(h->stream));
    const uint64_t base = h->shard_base;
    for (uint64_t k = 0; k < ncopy; ++k) {  // widen into the caller's (Go) memory
      due_idx[k] = base + hi[k];
      due_action[k] = ha[k];
    }
  }
  return n > cap ? AM_E_NOSPACE : AM_OK;
}

int am_sweep_tick_device(am_sweep_t* h, int64_t unix_sec, uint32_t mode, void* d_due_idx,
                         void* d_due_action, uint64_t cap, void* d_count, void* d_stats,
                         void* cuda_stream) {
  if (!h || (cap && (!d_due_idx || !d_due_action))) return AM_E_INVAL;
  if (unix_sec >= (1ll << 55) || unix_sec <= -(1ll << 55)) return AM_E_RANGE;
  TickGuard g(h);
  if (!g.ok) return AM_E_BUSY;
  AM_CUDA(h, cudaSetDevice(h->device));
  int rc = drain_staged(h);
  if (rc != AM_OK) return rc;
  cudaStream_t s = (cudaStream_t)cuda_stream;  // NULL == CUDA default stream
  return launch_sweep(h, unix_sec, mode, (uint32_t*)d_due_idx, (uint8_t*)d_due_action, cap,
                      (am_tick_stats_t*)d_stats, (uint32_t*)d_count, s);
}

int am_sweep_run_ticks(am_sweep_t* h, int64_t unix_sec0, uint64_t n_ticks, uint32_t mode,
                       uint64_t seed, am_tick_stats_t* stats_out) {
  if (!h || !stats_out) return AM_E_INVAL;
  if (n_ticks == 0) return AM_OK;
  if (n_ticks > (1ull << 24)) return AM_E_INVAL;
  if (unix_sec0 >= (1ll << 55) - (int64_t)n_ticks || unix_sec0 <= -(1ll << 55)) return AM_E_RANGE;
  TickGuard g(h);
  if (!g.ok) return AM_E_BUSY;
  AM_CUDA(h, cudaSetDevice(h->device));
  int rc = drain_staged(h);
  if (rc != AM_OK) return rc;
  h->seed = seed;
  am_tick_stats_t* d_stats = nullptr;
  AM_CUDA(h, cudaMalloc((void**)&d_stats, n_ticks * sizeof(am_tick_stats_t)));
  if (cudaError_t e = cudaEventRecord(h->ev0, h->stream); e != cudaSuccess) {
    h->last_error = cudaGetErrorString(e);
    rc = AM_E_DEVICE;
  }
  for (uint64_t k = 0; k < n_ticks && rc == AM_OK; ++k) {
    rc = launch_sweep(h, unix_sec0 + (int64_t)k, mode, h->due_idx[k & 1], h->due_action[k & 1],
                      h->cap_padded, d_stats + k, nullptr, h->stream);
  }
  if (rc == AM_OK) {
    cudaError_t e = cudaEventRecord(h->ev1, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (e == cudaSuccess) e = cudaMemcpy(stats_out, d_stats, n_ticks * sizeof(am_tick_stats_t), cudaMemcpyDeviceToHost);
    float ms = 0;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, h->ev0, h->ev1);
    if (e != cudaSuccess) { h->last_error = cudaGetErrorString(e); rc = AM_E_DEVICE; }
    else h->last_ms = ms;
  }
  cudaFree(d_stats);
  return rc;
}

int am_sweep_repeat_after_sec(am_sweep_t* h, int64_t unix_sec, uint64_t first, uint64_t n, int64_t* out) {
  if (!h || (n && !out)) return AM_E_INVAL;
  if (first + n > h->capacity || first + n < first || n > 0xFFFFFFFFull) return AM_E_INVAL;
  if (unix_sec >= (1ll << 55) || unix_sec <= -(1ll << 55)) return AM_E_RANGE;
  if (n == 0) return AM_OK;
  TickGuard g(h);
  if (!g.ok) return AM_E_BUSY;
  AM_CUDA(h, cudaSetDevice(h->device));
  int rc = drain_staged(h);
  if (rc != AM_OK) return rc;
  AM_CUDA(h, h->dev_in.reserve(n * 8));
  AM_CUDA(h, h->pin_out.reserve(n * 8));
  AM_LAUNCH(next_fire_kernel, (unsigned)((n + 127) / 128), 128, h->stream, h->cols, (uint32_t)first, (uint32_t)n,
            unix_sec, (int64_t*)h->dev_in.p);
  h->launches++;
  AM_CUDA(h, cudaGetLastError());
  AM_CUDA(h, cudaMemcpyAsync(h->pin_out.p, h->dev_in.p, n * 8, cudaMemcpyDeviceToHost, h->stream));
  AM_CUDA(h, cudaStreamSynchronize(h->stream));
  memcpy(out, h->pin_out.p, n * 8);
  return AM_OK;
}

int am_sweep_read(am_sweep_t* h, uint64_t first, uint64_t n, const uint64_t* idx,
                  am_record_cols_t* out) {
  if (!h || !out) return AM_E_INVAL;
  if (n == 0) return AM_OK;
  TickGuard g(h);
  if (!g.ok) return AM_E_BUSY;
  AM_CUDA(h, cudaSetDevice(h->device));
  {  // a read observes every upsert / remove / result staged before it
    int rc = drain_staged(h);
    if (rc != AM_OK) return rc;
  }
  if (!idx) {
    if (first + n > h->capacity || first + n < first) return AM_E_INVAL;
    for (int k = 0; k < 16; ++k) {
      void* dst = *cols_member(out, k);
      if (!dst) continue;
      AM_CUDA(h, cudaMemcpyAsync(dst, (char*)h->col_ptr[k] + first * kColElem[k], n * kColElem[k],
                                 cudaMemcpyDeviceToHost, h->stream));
    }
    AM_CUDA(h, cudaStreamSynchronize(h->stream));
    return AM_OK;
  }
  if (n > 0xFFFFFFFFull) return AM_E_INVAL;
  for (uint64_t k = 0; k < n; ++k)
    if (idx[k] >= h->capacity) return AM_E_RANGE;
  const size_t o_idx = n * sizeof(am_record_t), total = o_idx + n * 4;
  AM_CUDA(h, h->pin_in.reserve(total));
  AM_CUDA(h, h->dev_in.reserve(total));
  uint32_t* hidx = (uint32_t*)((char*)h->pin_in.p + o_idx);
  for (uint64_t k = 0; k < n; ++k) hidx[k] = (uint32_t)idx[k];
  char* dp = (char*)h->dev_in.p;
  AM_CUDA(h, cudaMemcpyAsync(dp + o_idx, hidx, n * 4, cudaMemcpyHostToDevice, h->stream));
  AM_LAUNCH(gather_records_kernel, (unsigned)((n + 255) / 256), 256, h->stream, h->cols,
            (const uint32_t*)(dp + o_idx), (am_record_t*)dp, (uint32_t)n);
  h->launches++;
  AM_CUDA(h, cudaGetLastError());
  AM_CUDA(h, cudaMemcpyAsync(h->pin_in.p, dp, n * sizeof(am_record_t), cudaMemcpyDeviceToHost, h->stream));
  AM_CUDA(h, cudaStreamSynchronize(h->stream));
  const am_record_t* r = (const am_record_t*)h->pin_in.p;
  for (uint64_t k = 0; k < n; ++k) {
    if (out->minute) out->minute[k] = r[k].minute;
    if (out->hour) out->hour[k] = r[k].hour;
    if (out->dom) out->dom[k] = r[k].dom;
    if (out->month) out->month[k] = r[k].month;
    if (out->dow) out->dow[k] = r[k].dow;
    if (out->ras) out->ras[k] = r[k].ras;
    if (out->flags) out->flags[k] = r[k].flags;
    if (out->finished_at) out->finished_at[k] = r[k].finished_at;
    if (out->runs_limit) out->runs_limit[k] = r[k].runs_limit;
    if (out->reset_interval) out->reset_interval[k] = r[k].reset_interval;
    if (out->success) out->success[k] = r[k].success;
    if (out->failed) out->failed[k] = r[k].failed;
    if (out->remedy_success) out->remedy_success[k] = r[k].remedy_success;
    if (out->remedy_failed) out->remedy_failed[k] = r[k].remedy_failed;
    if (out->remedy_total) out->remedy_total[k] = r[k].remedy_total;
    if (out->remedy_finished_at) out->remedy_finished_at[k] = r[k].remedy_finished_at;
  }
  return AM_OK;
}

}  // extern "C"
