// emu_sweep.cpp — the per-tick kernels of csrc/sweep_kernels.cuh (sweep_tick_kernel in its
// four variants, compact_kernel, publish_kernel, the staged-event kernels, next_fire_kernel,
// gather_records_kernel) compiled for the CPU emulator (cuda_emu.h) behind a small C API, so
// that tests/test_sweep_emulator.py can hold the SAME kernel source the GPU runs against the
// oracle without a GPU.  The launch sequence mirrors launch_sweep() / drain_staged() of
// csrc/sweep.cu.  Test infrastructure only.
#include "cuda_emu.h"

#include <cstdio>
#include <cstring>

#include "../../active-monitor_b200/csrc/sweep_kernels.cuh"

using namespace amsweep;

namespace {
constexpr size_t kColElem[16] = {8, 8, 8, 8, 8, 4, 4, 8, 4, 4, 4, 4, 4, 4, 4, 8};

struct EmuSweep {
  uint64_t capacity = 0, cap_padded = 0, shard_base = 0, n_records = 0;
  void* col[16] = {};
  DevCols cols{};
  uint32_t* seg_idx = nullptr;
  uint8_t* seg_act = nullptr;
  uint32_t* tile_count = nullptr;
  uint32_t* group_count[2] = {nullptr, nullptr};
  unsigned long long* acc = nullptr;
  uint32_t* marks = nullptr;
  uint32_t parity = 0;
};

void* zalloc(size_t bytes) {
  void* p = std::aligned_alloc(256, (bytes + 255) / 256 * 256);
  std::memset(p, 0, (bytes + 255) / 256 * 256);
  return p;
}
}  // namespace

// The product library exports host stubs with the same mangled names as the kernels compiled
// here (amsweep::sweep_tick_kernel<...>, ...), and the tests load it RTLD_GLOBAL in the same
// process: this library is therefore linked -Bsymbolic with hidden visibility (emu_sweep.py),
// only the C API below is exported.
#define EMU_API __attribute__((visibility("default")))
extern "C" {
EMU_API void* emu_sweep_create(uint64_t, uint64_t);
EMU_API void emu_sweep_destroy(void*);
EMU_API int emu_sweep_load(void*, uint64_t, uint64_t, const void* const*);
EMU_API int emu_sweep_read(void*, uint64_t, uint64_t, void* const*);
EMU_API int emu_sweep_tick(void*, int64_t, uint32_t, uint64_t, uint32_t*, uint8_t*, uint64_t, uint32_t*, am_tick_stats_t*);
EMU_API int emu_sweep_apply_ops(void*, const StagedOp*, uint32_t, const am_record_t*, int, int);
EMU_API int emu_sweep_repeat_after_sec(void*, int64_t, uint32_t, uint32_t, int64_t*);
EMU_API int emu_sweep_gather(void*, const uint32_t*, am_record_t*, uint32_t);
EMU_API uint32_t emu_op_upsert(void);
EMU_API uint32_t emu_op_remove(void);
EMU_API uint32_t emu_op_result(void);

void* emu_sweep_create(uint64_t capacity, uint64_t shard_base) {
  EmuSweep* h = new EmuSweep();
  h->capacity = capacity;
  h->cap_padded = (capacity + kTile - 1) / kTile * kTile;
  h->shard_base = shard_base;
  for (int k = 0; k < 16; ++k) h->col[k] = zalloc(h->cap_padded * kColElem[k]);
  void** slots[16] = {(void**)&h->cols.minute, (void**)&h->cols.hour, (void**)&h->cols.dom, (void**)&h->cols.month,
                      (void**)&h->cols.dow, (void**)&h->cols.ras, (void**)&h->cols.flags, (void**)&h->cols.finished_at,
                      (void**)&h->cols.runs_limit, (void**)&h->cols.reset_interval, (void**)&h->cols.success,
                      (void**)&h->cols.failed, (void**)&h->cols.remedy_success, (void**)&h->cols.remedy_failed,
                      (void**)&h->cols.remedy_total, (void**)&h->cols.remedy_finished_at};
  for (int k = 0; k < 16; ++k) *slots[k] = h->col[k];
  emu::launch(fill_u32_kernel, dim3(3), dim3(256), h->cols.flags, (uint32_t)AM_F_TOMBSTONE, h->cap_padded);
  const size_t ntiles = h->cap_padded / kTile, ngroups = (ntiles + kGroupTiles - 1) / kGroupTiles;
  h->seg_idx = (uint32_t*)zalloc(h->cap_padded * 4);
  h->seg_act = (uint8_t*)zalloc(h->cap_padded);
  h->tile_count = (uint32_t*)zalloc(ntiles * 4);
  for (int b = 0; b < 2; ++b) h->group_count[b] = (uint32_t*)zalloc(ngroups * 4);
  h->acc = (unsigned long long*)zalloc(kNumAcc * 8);
  h->marks = (uint32_t*)zalloc(h->cap_padded * 8);
  return h;
}

void emu_sweep_destroy(void* hv) {
  EmuSweep* h = (EmuSweep*)hv;
  if (!h) return;
  for (int k = 0; k < 16; ++k) std::free(h->col[k]);
  std::free(h->seg_idx); std::free(h->seg_act); std::free(h->tile_count);
  std::free(h->group_count[0]); std::free(h->group_count[1]); std::free(h->acc); std::free(h->marks);
  delete h;
}

// am_sweep_load_range
int emu_sweep_load(void* hv, uint64_t first, uint64_t n, const void* const* cols) {
  EmuSweep* h = (EmuSweep*)hv;
  if (first + n > h->capacity) return -1;
  for (int k = 0; k < 16; ++k) {
    char* dst = (char*)h->col[k] + first * kColElem[k];
    if (cols[k]) std::memcpy(dst, cols[k], n * kColElem[k]);
    else std::memset(dst, 0, n * kColElem[k]);
  }
  if (first + n > h->n_records) h->n_records = first + n;
  return 0;
}

int emu_sweep_read(void* hv, uint64_t first, uint64_t n, void* const* cols) {
  EmuSweep* h = (EmuSweep*)hv;
  if (first + n > h->capacity) return -1;
  for (int k = 0; k < 16; ++k)
    if (cols[k]) std::memcpy(cols[k], (char*)h->col[k] + first * kColElem[k], n * kColElem[k]);
  return 0;
}

// launch_sweep() of csrc/sweep.cu, kernel for kernel
int emu_sweep_tick(void* hv, int64_t T, uint32_t mode, uint64_t seed, uint32_t* out_idx, uint8_t* out_act,
                   uint64_t cap, uint32_t* out_count, am_tick_stats_t* out_stats) {
  EmuSweep* h = (EmuSweep*)hv;
  if (h->n_records == 0) {
    if (out_stats) std::memset(out_stats, 0, sizeof *out_stats);
    if (out_count) *out_count = 0;
    return 0;
  }
  SweepParams p{};
  p.c = h->cols;
  p.n_records = h->n_records;
  p.shard_base = h->shard_base;
  p.seed = seed;
  p.T = T;
  p.words = tick_words_from_unix(T);
  p.n_tiles = (uint32_t)((h->n_records + kTile - 1) / kTile);
  p.mode = mode;
  p.seg_idx = h->seg_idx;
  p.seg_act = h->seg_act;
  p.tile_count = h->tile_count;
  p.group_count = h->group_count[h->parity];
  p.acc = h->acc;
  int64_t sec_of_min = T % 60;
  if (sec_of_min < 0) sec_of_min += 60;
  const bool masks = sec_of_min == 0 || (mode & AM_SWEEP_FULL_SCAN);
  const bool closed = (mode & AM_SWEEP_CLOSED_LOOP) != 0;
  const dim3 grid(p.n_tiles), block(kBlock);
  if (closed && masks) emu::launch(sweep_tick_kernel<true, true>, grid, block, p);
  else if (closed) emu::launch(sweep_tick_kernel<true, false>, grid, block, p);
  else if (masks) emu::launch(sweep_tick_kernel<false, true>, grid, block, p);
  else emu::launch(sweep_tick_kernel<false, false>, grid, block, p);
  CompactParams c{};
  c.seg_idx = h->seg_idx;
  c.seg_act = h->seg_act;
  c.tile_count = h->tile_count;
  c.group_count = h->group_count[h->parity];
  c.group_count_next = h->group_count[h->parity ^ 1];
  c.acc = h->acc;
  c.out_idx = out_idx;
  c.out_act = out_act;
  c.shard_base = h->shard_base;
  c.n_tiles = p.n_tiles;
  c.n_groups = (p.n_tiles + kGroupTiles - 1) / kGroupTiles;
  c.cap = (uint32_t)(cap > 0xFFFFFFFFull ? 0xFFFFFFFFull : cap);
  emu::launch(compact_kernel, dim3(c.n_groups), dim3(256), c);
  emu::launch(publish_kernel, dim3(1), dim3(32), h->acc, out_stats, out_count, h->n_records);
  h->parity ^= 1;
  return 0;
}

// drain_staged() of csrc/sweep.cu: ops in call order, records of the upserts in `recs`
int emu_sweep_apply_ops(void* hv, const StagedOp* ops, uint32_t n, const am_record_t* recs, int n_state, int n_result) {
  EmuSweep* h = (EmuSweep*)hv;
  if (n == 0) return 0;
  if (n_state)
    for (uint32_t k = 0; k < n; ++k)
      if ((ops[k].arg & kOpKindMask) == kOpUpsert && (uint64_t)ops[k].idx + 1 > h->n_records)
        h->n_records = (uint64_t)ops[k].idx + 1;
  const dim3 grid((n + 255) / 256), block(256);
  emu::launch(mark_ops_kernel, grid, block, h->marks, ops, n);
  if (n_state) emu::launch(apply_state_ops_kernel, grid, block, h->cols, (const uint32_t*)h->marks, ops, recs, n);
  if (n_result) emu::launch(apply_result_ops_kernel, grid, block, h->cols.flags, (const uint32_t*)h->marks, ops, n);
  emu::launch(clear_marks_kernel, grid, block, h->marks, ops, n);
  for (uint64_t i = 0; i < 2 * h->cap_padded; ++i)
    if (h->marks[i]) return -2;  // marks must be back to zero
  return 0;
}

int emu_sweep_repeat_after_sec(void* hv, int64_t T, uint32_t first, uint32_t n, int64_t* out) {
  EmuSweep* h = (EmuSweep*)hv;
  emu::launch(next_fire_kernel, dim3((n + 127) / 128), dim3(128), h->cols, first, n, T, out);
  return 0;
}

int emu_sweep_gather(void* hv, const uint32_t* idx, am_record_t* out, uint32_t n) {
  EmuSweep* h = (EmuSweep*)hv;
  emu::launch(gather_records_kernel, dim3((n + 255) / 256), dim3(256), h->cols, idx, out, n);
  return 0;
}

uint32_t emu_op_upsert(void) { return kOpUpsert; }
uint32_t emu_op_remove(void) { return kOpRemove; }
uint32_t emu_op_result(void) { return kOpResult; }

}  // extern "C"
