// Cross-rank ordering of the tick exchange under ThreadSanitizer: three ranks (threads) each own a
// shard (am_sweep on its own emulated device), tick and exchange back to back through the C-ABI —
// am_sweep_tick_shard + am_gather_exchange — with no barrier between ticks.  Built from the
// library's own sources on the emulator with -fsanitize=thread.  System-scope release/acquire are
// atomics; peer payload stores and the list rebuild's reads are plain accesses, so any TSan report
// means cross-rank data that the done-flag protocol does not order.  Every rank also checks that
// all ranks rebuilt the SAME global list (checksum), ascending.
#include <atomic>
#include <barrier>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/amsweep.h"

namespace {
constexpr int kWorld = 3, kTicks = 6;
constexpr uint64_t kPerRank = 20000;
char g_handles[kWorld][AM_IPC_HANDLE_BYTES];
std::atomic<unsigned long long> g_sum[kTicks][kWorld];
std::atomic<int> g_fail{0};
}  // namespace

int main() {
  std::barrier bar(kWorld);
  std::vector<std::thread> ranks;
  for (int r = 0; r < kWorld; ++r)
    ranks.emplace_back([&, r] {
      const uint64_t base = (uint64_t)r * kPerRank, n = kPerRank + (r == kWorld - 1 ? 37 : 0);
      am_sweep_t* s = nullptr;
      am_gather_t* g = nullptr;
      if (am_sweep_create(&s, r, n, base) != AM_OK) { g_fail = 1; return; }
      // a mixed population: interval checks with different phases, some paused, some with results posted
      std::vector<uint64_t> idx(n);
      std::vector<am_record_t> recs(n);
      for (uint64_t i = 0; i < n; ++i) {
        am_healthcheck_t hc;
        std::memset(&hc, 0, sizeof hc);
        hc.has_resource = 1;
        hc.timer_armed = (i % 17) != 0;
        hc.repeat_after_sec = (i % 11 == 0) ? 0 : (int64_t)(5 + (i * 7 + r) % 90);
        hc.finished_at_set = 1;
        hc.finished_at = 1789982100 - (int64_t)((i * 13 + r * 5) % 120);
        hc.has_remedy = i % 3 == 0;
        idx[i] = i;
        if (am_healthcheck_classify(&hc, &recs[i]) != AM_OK) g_fail = 1;
      }
      if (am_sweep_upsert(s, n, idx.data(), recs.data()) != AM_OK) g_fail = 1;
      const uint64_t cap_total = (uint64_t)kWorld * kPerRank + 37;
      if (am_gather_create(&g, r, r, kWorld, cap_total, 4) != AM_OK) { g_fail = 1; return; }
      if (am_gather_export(g, g_handles[r]) != AM_OK) g_fail = 1;
      bar.arrive_and_wait();
      if (am_gather_connect(g, g_handles) != AM_OK) g_fail = 1;
      uint64_t bases[kWorld], sizes[kWorld];
      for (int q = 0; q < kWorld; ++q) { bases[q] = (uint64_t)q * kPerRank; sizes[q] = kPerRank + (q == kWorld - 1 ? 37 : 0); }
      if (am_gather_set_layout(g, bases, sizes) != AM_OK) g_fail = 1;
      for (int k = 0; k < kTicks && !g_fail; ++k) {
        if (k == 2) {  // results posted between ticks: exceptions (remedy actions) on the wire
          std::vector<uint64_t> pi;
          std::vector<uint8_t> ph;
          for (uint64_t i = 0; i < n; i += 5) { pi.push_back(i); ph.push_back((uint8_t)(1 + i % 2)); }
          if (am_sweep_post_result(s, pi.size(), pi.data(), ph.data(), nullptr) != AM_OK) g_fail = 1;
        }
        am_tick_stats_t st;
        if (am_sweep_tick_shard(s, 1789982100 + 7 * k, 0, nullptr) != AM_OK) g_fail = 1;
        if (am_gather_exchange(g, s, &st, nullptr) != AM_OK) g_fail = 1;
        const uint32_t* counts = (const uint32_t*)am_gather_out_counts(g);
        const uint32_t total = counts[kWorld];
        const uint32_t* gi = (const uint32_t*)am_gather_out_idx(g);
        const uint8_t* ga = (const uint8_t*)am_gather_out_act(g);
        unsigned long long sum = total;
        for (uint32_t e = 0; e < total; ++e) {
          if (e && gi[e] <= gi[e - 1]) g_fail = 1;
          sum = sum * 1099511628211ull + gi[e] * 131ull + ga[e];
        }
        if (st.n_emitted != counts[r]) g_fail = 1;
        g_sum[k][r] = sum;
      }
      bar.arrive_and_wait();
      am_gather_destroy(g);
      am_sweep_destroy(s);
    });
  for (auto& t : ranks) t.join();
  for (int k = 0; k < kTicks; ++k)
    for (int r = 1; r < kWorld; ++r)
      if (g_sum[k][r] != g_sum[k][0]) g_fail = 1;
  if (g_fail) { std::printf("FAILED\n"); return 1; }
  std::printf("ok %d ranks x %d ticks\n", kWorld, kTicks);
  return 0;
}
