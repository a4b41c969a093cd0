// Thread-safety of the C-ABI's threading contract: upsert / remove / post_result
// from several threads (the reference's <= MaxParallel Reconcile workers and watch goroutines),
// ONE ticker thread calling am_sweep_tick / am_sweep_read.  Built with -fsanitize=thread
// together with the emulated library sources; ThreadSanitizer checks the host runtime's locking
// (staging arrays, high-water mark, tick guard).  Exit code 0 and no TSan report = pass.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/amsweep.h"

int main() {
  const uint64_t cap = 4096;
  am_sweep_t* h = nullptr;
  if (am_sweep_create(&h, 0, cap, 0) != AM_OK) { std::printf("create failed\n"); return 1; }
  am_healthcheck_t hc;
  std::memset(&hc, 0, sizeof hc);
  hc.repeat_after_sec = 5;
  hc.has_resource = 1;
  hc.finished_at_set = 1;
  hc.finished_at = 1789982000;
  am_record_t rec;
  if (am_healthcheck_classify(&hc, &rec) != AM_OK) return 1;
  std::atomic<bool> stop{false};
  std::atomic<uint64_t> calls{0};
  std::vector<std::thread> producers;
  for (int t = 0; t < 4; ++t)
    producers.emplace_back([&, t] {
      uint64_t x = 88172645463325252ull + (uint64_t)t;
      for (int it = 0; it < 1500 && !stop.load(); ++it) {
        std::this_thread::sleep_for(std::chrono::microseconds(100));  // a flood would only test the emulator's speed
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        uint64_t idx[3] = {x % cap, (x >> 20) % cap, (x >> 40) % cap};
        am_record_t recs[3] = {rec, rec, rec};
        uint8_t ph[3] = {(uint8_t)(x % 3), (uint8_t)((x >> 8) % 3), 1}, rp[3] = {0, (uint8_t)((x >> 16) % 3), 2};
        switch ((x >> 50) % 4) {
          case 0: am_sweep_upsert(h, 3, idx, recs); break;
          case 1: am_sweep_post_result(h, 3, idx, ph, rp); break;
          case 2: am_sweep_post_result(h, 2, idx, ph, nullptr); break;
          default: am_sweep_remove(h, 1, idx); break;
        }
        calls.fetch_add(1);
      }
    });
  std::vector<uint64_t> di(cap);
  std::vector<uint32_t> da(cap);
  uint64_t total = 0;
  for (int k = 0; k < 150; ++k) {
    uint64_t n = 0;
    am_tick_stats_t st;
    const int rc = am_sweep_tick(h, 1789982100 + k, k % 3 == 0 ? AM_SWEEP_FULL_SCAN : 0u, di.data(), da.data(), cap, &n, &st);
    if (rc != AM_OK) { std::printf("tick rc %d\n", rc); return 1; }
    if (st.n_emitted != n || st.n_records > cap) { std::printf("inconsistent stats\n"); return 1; }
    for (uint64_t i = 1; i < n; ++i)
      if (di[i] <= di[i - 1]) { std::printf("list not ascending\n"); return 1; }
    total += n;
    if (k % 10 == 0) {
      uint64_t q[4] = {1, 77, 1024, cap - 1};
      am_record_t out[4];
      am_record_cols_t c;
      std::memset(&c, 0, sizeof c);
      uint32_t flags[4];
      c.flags = flags;
      (void)out;
      if (am_sweep_read(h, 0, 4, q, &c) != AM_OK) { std::printf("read failed\n"); return 1; }
    }
  }
  stop.store(true);
  for (auto& t : producers) t.join();
  am_sweep_destroy(h);
  std::printf("ok %llu staged calls, %llu entries emitted\n", (unsigned long long)calls.load(), (unsigned long long)total);
  return 0;
}
