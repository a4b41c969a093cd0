// Host-only entry points of the C-ABI called the way a compiled controller would call them
// (plain C++ threads, no Python, no GPU): bulk ingest (am_healthcheck_classify_batch) and the
// hand-off queue between the ticker and the MaxParallel workers (am_handoff_*).
//
// Mirrors what the reference guarantees at these seams: processHealthCheck's ladder per CR
// (hcc.go:227/238/251/264) whatever the batching, and exactly one createSubmitWorkflow per due
// check (one AfterFunc fire per timer, hcc.go:751) with bounded concurrency (hcc.go:298).
#include <atomic>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "amsweep.h"

#define CHECK(cond)                                                         \
  do {                                                                      \
    if (!(cond)) {                                                          \
      std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);         \
      return 1;                                                             \
    }                                                                       \
  } while (0)

static int test_classify_batch() {
  const char* crons[] = {"", "@every 5s", "*/5 * * * *", "NOT_A_VALID_CRON", "0 9 * * mon-fri",
                         "CRON_TZ=Asia/Tokyo 0 12 * * *", "@hourly", "60 * * * *"};
  const uint64_t n = 50000;
  std::vector<am_healthcheck_t> in(n);
  for (uint64_t i = 0; i < n; ++i) {
    am_healthcheck_t& h = in[i];
    std::memset(&h, 0, sizeof h);
    h.cron = crons[i % 8];
    h.cron_len = std::strlen(h.cron);
    h.repeat_after_sec = (i % 5 == 0) ? 60 : (i % 7 == 0 ? -1 : 0);
    h.has_resource = i % 50 != 0;
    h.has_remedy = i % 3 == 0;
    h.remedy_runs_limit = i % 4;
    h.remedy_reset_interval = (i % 2) * 300;
    if (i % 9) { h.finished_at = 1789982100 - (int64_t)(i % 7200); h.finished_at_set = 1; }
    h.success_count = (int64_t)(i % 1000);
    if (i % 1013 == 0) h.failed_count = 1ll << 40;  // does not fit the column: AM_E_RANGE
  }
  std::vector<am_record_t> one(n), many(n);
  std::vector<int32_t> rc1(n), rcn(n);
  uint64_t bad1 = 0, badn = 0;
  CHECK(am_healthcheck_classify_batch(in.data(), n, one.data(), rc1.data(), 1, &bad1) == AM_OK);
  CHECK(am_healthcheck_classify_batch(in.data(), n, many.data(), rcn.data(), 6, &badn) == AM_OK);
  CHECK(bad1 == badn && bad1 > 0);
  CHECK(std::memcmp(one.data(), many.data(), n * sizeof(am_record_t)) == 0);
  CHECK(rc1 == rcn);
  uint64_t bad = 0;
  for (uint64_t i = 0; i < n; ++i) {
    am_record_t r;
    const int rc = am_healthcheck_classify(&in[i], &r);
    CHECK(rc == rc1[i]);
    CHECK(std::memcmp(&r, &one[i], sizeof r) == 0);
    bad += rc != AM_OK;
    if (rc == AM_OK && (i % 50 != 0) && in[i].repeat_after_sec <= 0 && i % 8 == 3)
      CHECK((r.flags & AM_KIND_MASK) == AM_KIND_PARSE_ERROR);  // unit_test.go:617-634
    if (rc == AM_OK && (i % 50 != 0) && in[i].repeat_after_sec <= 0 && i % 8 == 1)
      CHECK((r.flags & AM_KIND_MASK) == AM_KIND_CRON_EVERY && r.ras == 5);  // unit_test.go:636-660
    if (rc == AM_OK && (i % 50 != 0) && in[i].repeat_after_sec <= 0 && i % 8 == 0)
      CHECK((r.flags & AM_KIND_MASK) == AM_KIND_STOPPED);  // healthcheck_controller_test.go:119-156
  }
  CHECK(bad == bad1);
  CHECK(am_healthcheck_classify_batch(nullptr, 0, nullptr, nullptr, 0, nullptr) == AM_OK);
  CHECK(am_healthcheck_classify_batch(nullptr, 3, nullptr, nullptr, 0, nullptr) == AM_E_INVAL);
  return 0;
}

static int test_handoff() {
  am_handoff_t* q = nullptr;
  CHECK(am_handoff_create(&q, 4096) == AM_OK);
  const int workers = 6, ticks = 400, per_tick = 700;
  std::vector<std::vector<am_work_item_t>> seen(workers);
  std::atomic<bool> done{false};
  std::vector<std::thread> th;
  for (int w = 0; w < workers; ++w)
    th.emplace_back([&, w] {
      am_work_item_t buf[97];
      for (;;) {
        uint64_t n = 0;
        if (am_handoff_pop(q, 97, buf, &n) != AM_OK) return;
        if (n) {
          seen[w].insert(seen[w].end(), buf, buf + n);
          continue;
        }
        uint64_t pending = 1;
        am_handoff_stats(q, &pending, nullptr, nullptr, nullptr);
        if (done.load() && pending == 0) return;
        std::this_thread::yield();
      }
    });
  std::vector<uint64_t> idx(per_tick);
  std::vector<uint32_t> act(per_tick);
  uint64_t expected = 0, rejected_seen = 0;
  for (int t = 0; t < ticks; ++t) {
    uint64_t want = 0;
    for (int k = 0; k < per_tick; ++k) {
      idx[k] = (uint64_t)t * per_tick + k;
      // a third submits, a sixth also runs a remedy, the rest are ticker-only actions
      act[k] = k % 3 == 0 ? (AM_ACT_SUBMIT_HC | (k % 6 == 0 ? AM_ACT_RUN_REMEDY : 0u)) : AM_ACT_STOPPED;
      want += k % 3 == 0;
    }
    for (;;) {  // back-pressure: retry the whole tick
      uint64_t n = 0;
      const int rc = am_handoff_publish(q, 1000 + t, per_tick, idx.data(), act.data(),
                                        AM_ACT_SUBMIT_HC | AM_ACT_RUN_REMEDY, &n);
      CHECK(n == want);
      if (rc == AM_OK) break;
      CHECK(rc == AM_E_NOSPACE);
      ++rejected_seen;
      std::this_thread::yield();
    }
    expected += want;
  }
  done.store(true);
  for (auto& t : th) t.join();
  uint64_t total = 0;
  std::vector<uint8_t> hit((size_t)ticks * per_tick, 0);
  for (auto& v : seen) {
    total += v.size();
    for (size_t k = 0; k < v.size(); ++k) {
      if (k) CHECK(v[k].idx > v[k - 1].idx);  // every worker sees a subsequence of the global order
      CHECK(v[k].idx % per_tick % 3 == 0);
      CHECK(v[k].unix_sec == 1000 + (int64_t)(v[k].idx / per_tick));
      CHECK((v[k].action & ~(AM_ACT_SUBMIT_HC | AM_ACT_RUN_REMEDY)) == 0 && (v[k].action & AM_ACT_SUBMIT_HC));
      CHECK(hit[v[k].idx]++ == 0);  // exactly one worker per due check
    }
  }
  CHECK(total == expected);
  uint64_t pending = 9, published = 0, popped = 0, rejected = 0;
  CHECK(am_handoff_stats(q, &pending, &published, &popped, &rejected) == AM_OK);
  CHECK(pending == 0 && published == expected && popped == expected && rejected == rejected_seen);
  am_handoff_destroy(q);
  return 0;
}

int main() {
  if (test_classify_batch()) return 1;
  if (test_handoff()) return 1;
  std::printf("ok\n");
  return 0;
}
