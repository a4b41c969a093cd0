// cuda_emu.h — just enough of the CUDA execution model to run csrc/gather_kernels.cuh on a
// CPU: test infrastructure, never part of the product.
//
//   * one OS thread per rank ("GPU"); the CTAs of a launch run one after the other on it;
//   * the 256 threads of a CTA are fibers (own stacks, a minimal register switch), scheduled round-robin; __syncthreads and
//     the warp collectives (__shfl_up_sync, __reduce_add_sync) are rendezvous points;
//   * __shared__ is static thread_local (one copy per rank thread, reused by successive
//     CTAs, as on an SM); system-scope acquire/release and atomics are seq_cst std atomics;
//     a spinning fiber yields to its siblings.
//
// What it checks: index arithmetic, buffer offsets, the epoch / ticket / done-flag protocol
// and its double buffering under real concurrency between ranks.  What it cannot check:
// the GPU memory model beyond x86-TSO, coalescing, occupancy, performance.
#pragma once
#include <stdint.h>
#include <string.h>
#if !defined(__x86_64__)
#include <ucontext.h>
#endif

#include <atomic>
#include <chrono>
#include <cstdlib>
#include <functional>
#include <random>
#include <thread>
#include <vector>

#define AMSWEEP_EMULATE 1
#define __global__
#define __device__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static thread_local
#define __launch_bounds__(...)
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { uint32_t x, y; } __attribute__((aligned(8)));
struct uint4 { uint32_t x, y, z, w; } __attribute__((aligned(16)));
struct ulonglong2 { unsigned long long x, y; } __attribute__((aligned(16)));
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
struct int2 { int32_t x, y; } __attribute__((aligned(8)));
struct int4 { int32_t x, y, z, w; } __attribute__((aligned(16)));
struct longlong2 { long long x, y; } __attribute__((aligned(16)));
static inline int2 make_int2(int32_t x, int32_t y) { return int2{x, y}; }
static inline longlong2 make_longlong2(long long x, long long y) { return longlong2{x, y}; }
struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };

#if defined(__SANITIZE_THREAD__)
// ThreadSanitizer must be told about the hand-made stack switches.  Every switch synchronises (the
// default), so the fibers of one rank thread are totally ordered and only accesses of DIFFERENT
// ranks can race: exactly what the flag protocol of the exchange has to order.
extern "C" {
void* __tsan_get_current_fiber(void);
void* __tsan_create_fiber(unsigned flags);
void __tsan_destroy_fiber(void* fiber);
void __tsan_switch_to_fiber(void* fiber, unsigned flags);
}
#define EMU_TSAN 1
#endif

namespace emu {
enum State { READY, WAIT_BLOCK, WAIT_WARP, DONE };
#if defined(__x86_64__)
// Minimal stack switch (callee-saved registers + rsp): glibc's swapcontext makes a sigprocmask
// system call per switch, and a CTA's 256 fibers switch at every rendezvous.
extern "C" void emu_switch(void** save_sp, void* load_sp);
asm(R"(
.text
.weak emu_switch
.hidden emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");
struct Context { void* sp = nullptr; };
#else
struct Context { ucontext_t uc; };
#endif
struct Fiber {
  Context ctx;
  void* tsan = nullptr;
  State state = DONE;
  dim3 tidx;
  unsigned long long slot = 0;  // value deposited for a warp collective
  char* stack = nullptr;
};
struct Block {
  std::vector<Fiber> fibers;
  Context sched;
  void* sched_tsan = nullptr;
  Fiber* cur = nullptr;
  dim3 bidx, bdim, gdim;
  void (*entry)(const void*) = nullptr;
  const void* params = nullptr;
};
inline thread_local Block* blk = nullptr;
constexpr size_t kStack = 256 * 1024;

#if defined(__x86_64__)
inline void switch_to(Context& from, Context& to) { emu_switch(&from.sp, to.sp); }
#else
inline void switch_to(Context& from, Context& to) { swapcontext(&from.uc, &to.uc); }
#endif
inline void to_sched() {
#ifdef EMU_TSAN
  __tsan_switch_to_fiber(blk->sched_tsan, 0);
#endif
  switch_to(blk->cur->ctx, blk->sched);
}
inline void yield_spin() { to_sched(); }  // stays READY
inline void trampoline() {
  blk->entry(blk->params);
  blk->cur->state = DONE;
  to_sched();
}
inline void warp_rendezvous() { blk->cur->state = WAIT_WARP; to_sched(); }

// run one CTA to completion on the calling OS thread
inline void run_block(Block& b) {
  blk = &b;
  const unsigned n = b.bdim.x;
  for (unsigned t = 0; t < n; ++t) {
    Fiber& f = b.fibers[t];
    f.tidx = dim3(t, 0, 0);
    f.state = READY;
#ifdef EMU_TSAN
    if (f.tsan) __tsan_destroy_fiber(f.tsan);
    f.tsan = __tsan_create_fiber(0);
    b.sched_tsan = __tsan_get_current_fiber();
#endif
#if defined(__x86_64__)
    // first switch "returns" into trampoline with the stack aligned as after a call
    uintptr_t top = ((uintptr_t)f.stack + kStack) & ~(uintptr_t)15;
    void** sp = (void**)top;
    *--sp = nullptr;                 // return address of trampoline (never used)
    *--sp = (void*)&trampoline;      // popped by emu_switch's ret
    for (int r = 0; r < 6; ++r) *--sp = nullptr;  // rbp rbx r12 r13 r14 r15
    f.ctx.sp = sp;
#else
    getcontext(&f.ctx.uc);
    f.ctx.uc.uc_stack.ss_sp = f.stack;
    f.ctx.uc.uc_stack.ss_size = kStack;
    f.ctx.uc.uc_link = &b.sched.uc;
    makecontext(&f.ctx.uc, (void (*)())trampoline, 0);
#endif
  }
  unsigned done = 0;
  while (done < n) {
    done = 0;
    for (unsigned t = 0; t < n; ++t) {
      Fiber& f = b.fibers[t];
      if (f.state == READY) {
        b.cur = &f;
#ifdef EMU_TSAN
        __tsan_switch_to_fiber(f.tsan, 0);
#endif
        switch_to(b.sched, f.ctx);
      }
      done += f.state == DONE;
    }
    // __syncthreads: released when every live thread of the CTA has arrived
    unsigned waiting = 0, live = 0;
    for (unsigned t = 0; t < n; ++t) {
      live += b.fibers[t].state != DONE;
      waiting += b.fibers[t].state == WAIT_BLOCK;
    }
    if (live && waiting == live)
      for (unsigned t = 0; t < n; ++t)
        if (b.fibers[t].state == WAIT_BLOCK) b.fibers[t].state = READY;
    // warp collectives: released per warp
    for (unsigned w0 = 0; w0 < n; w0 += 32) {
      unsigned wl = 0, ww = 0;
      for (unsigned t = w0; t < w0 + 32 && t < n; ++t) {
        wl += b.fibers[t].state != DONE;
        ww += b.fibers[t].state == WAIT_WARP;
      }
      if (wl && ww == wl)
        for (unsigned t = w0; t < w0 + 32 && t < n; ++t)
          if (b.fibers[t].state == WAIT_WARP) b.fibers[t].state = READY;
    }
  }
  blk = nullptr;
}

// EMU_JITTER=1: random pauses at launches and at system-scope stores, different per rank thread, so
// that ranks drift apart by whole kernels — the skew the epoch / double-buffer protocol must survive.
inline void jitter() {
  static const bool on = std::getenv("EMU_JITTER") != nullptr;
  if (!on) return;
  static thread_local std::mt19937 rng((unsigned)std::hash<std::thread::id>()(std::this_thread::get_id()));
  const unsigned r = rng();
  if (r % 4 == 0) std::this_thread::sleep_for(std::chrono::microseconds(r % 1500));
}

// kernel<<<grid, block>>>(args...) on the calling rank thread
template <class K, class... A>
void launch(K kernel, dim3 grid, dim3 block, const A&... args) {
  static thread_local Block b;
  if (b.fibers.size() < block.x) {
    const size_t old = b.fibers.size();
    b.fibers.resize(block.x);
    for (size_t t = old; t < block.x; ++t) b.fibers[t].stack = (char*)std::malloc(kStack);
  }
  jitter();
  auto call = [&]() { kernel(args...); };
  b.entry = [](const void* c) { (*(const decltype(call)*)c)(); };
  b.params = &call;
  b.bdim = block;
  b.gdim = grid;
  for (unsigned by = 0; by < grid.y; ++by)
    for (unsigned bx = 0; bx < grid.x; ++bx) {
      b.bidx = dim3(bx, by, 0);
      run_block(b);
    }
}
}  // namespace emu

#define threadIdx (emu::blk->cur->tidx)
#define blockIdx (emu::blk->bidx)
#define blockDim (emu::blk->bdim)
#define gridDim (emu::blk->gdim)

static inline void __syncthreads() { emu::blk->cur->state = emu::WAIT_BLOCK; emu::to_sched(); }
static inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
static inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
template <class T> static inline T __ldcs(const T* p) { return *p; }
static inline int __popc(uint32_t v) { return __builtin_popcount(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicOr(uint32_t* p, uint32_t v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicAnd(uint32_t* p, uint32_t v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
static inline uint32_t atomicXor(uint32_t* p, uint32_t v) { return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST); }

static inline void st_release_sys(unsigned long long* p, unsigned long long v) {
  emu::jitter();
  __atomic_store_n(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned long long ld_acquire_sys(const unsigned long long* p) {
  const unsigned long long v = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  emu::yield_spin();  // callers poll in a loop: let the sibling fibers run, other ranks are OS threads
  return v;
}

template <class T> static inline void __stcs(T* p, T v) { *p = v; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned long long atomicXor(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_xor(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline uint32_t atomicMax(uint32_t* p, uint32_t v) {
  uint32_t old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}

static inline unsigned long long global_timer_ns() {
  return (unsigned long long)std::chrono::duration_cast<std::chrono::nanoseconds>(
             std::chrono::steady_clock::now().time_since_epoch()).count();
}

// full-mask warp collectives (every lane of the warp takes part, as in the kernels under test)
static inline uint32_t __shfl_up_sync(unsigned, uint32_t v, unsigned delta) {
  emu::Block& b = *emu::blk;
  const unsigned t = b.cur->tidx.x, lane = t & 31;
  b.cur->slot = v;
  emu::warp_rendezvous();
  const uint32_t r = lane >= delta ? (uint32_t)b.fibers[t - delta].slot : v;
  emu::warp_rendezvous();  // nobody overwrites a slot before every lane has read
  return r;
}
static inline uint32_t __reduce_add_sync(unsigned, uint32_t v) {
  emu::Block& b = *emu::blk;
  const unsigned t = b.cur->tidx.x, w0 = t & ~31u;
  b.cur->slot = v;
  emu::warp_rendezvous();
  uint32_t s = 0;
  for (unsigned k = w0; k < w0 + 32 && k < b.bdim.x; ++k) s += (uint32_t)b.fibers[k].slot;
  emu::warp_rendezvous();
  return s;
}

// generic full-warp exchange: every live lane deposits `v`, then `f(slots of the warp, lane)`
namespace emu {
template <class F>
static inline auto warp_collect(unsigned long long v, F f) {
  Block& b = *blk;
  const unsigned t = b.cur->tidx.x, w0 = t & ~31u;
  b.cur->slot = v;
  warp_rendezvous();
  unsigned long long vals[32];
  bool live[32];
  for (unsigned k = 0; k < 32; ++k) {
    const bool in = w0 + k < b.bdim.x;
    live[k] = in && b.fibers[w0 + k].state != DONE;
    vals[k] = in ? b.fibers[w0 + k].slot : 0;
  }
  auto r = f(vals, live, t & 31u);
  warp_rendezvous();
  return r;
}
}  // namespace emu
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { emu::warp_rendezvous(); }
static inline unsigned __ballot_sync(unsigned, int pred) {
  return emu::warp_collect(pred ? 1ull : 0ull, [](const unsigned long long* v, const bool* live, unsigned) {
    unsigned m = 0;
    for (unsigned k = 0; k < 32; ++k) m |= (live[k] && v[k]) ? (1u << k) : 0u;
    return m;
  });
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline uint32_t __reduce_xor_sync(unsigned, uint32_t x) {
  return emu::warp_collect(x, [](const unsigned long long* v, const bool* live, unsigned) {
    uint32_t s = 0;
    for (unsigned k = 0; k < 32; ++k) s ^= live[k] ? (uint32_t)v[k] : 0u;
    return s;
  });
}
static inline unsigned long long __shfl_xor_sync(unsigned, unsigned long long x, int lane_mask) {
  return emu::warp_collect(x, [lane_mask](const unsigned long long* v, const bool*, unsigned lane) {
    return v[(lane ^ (unsigned)lane_mask) & 31u];
  });
}
