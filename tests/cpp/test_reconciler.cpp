// C++ tests of the host-side mirror (include/amsweep_reconciler.hpp), written to read like
// the reference's own tests of this path:
//   internal/controllers/healthcheck_controller_unit_test.go:617-660  (cron parsing)
//   internal/controllers/healthcheck_controller_test.go:53-76, :119-156, :158-202 (remedy flow,
//       pause, pre-armed timer)
//   internal/controllers/healthcheck_controller_edge_test.go:47-74, :152-199 (nil resource, delete)
// Needs a CUDA device (run by tests/test_reconciler_cpp.py under -m gpu).
#include <cstdio>
#include <cstdlib>
#include <string>

#include "amsweep_reconciler.hpp"

using namespace amsweep;

static int failures = 0;
#define CHECK(cond)                                                          \
  do {                                                                       \
    if (!(cond)) { std::printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); ++failures; } \
  } while (0)

static HealthCheck newHC(const std::string& name, const std::string& ns) {
  HealthCheck hc;
  hc.Name = name;
  hc.Namespace = ns;
  return hc;
}
static const int64_t T0 = 1789982100;  // 2026-09-21 09:15:00 UTC

static void TestProcessHealthCheck_InvalidCron_ReturnsError(HealthCheckReconciler& r) {
  HealthCheck hc = newHC("invalid-cron", "default");
  hc.Spec.RepeatAfterSec = 0;
  hc.Spec.Schedule.Cron = "NOT_A_VALID_CRON";
  auto err = r.ProcessHealthCheck(&hc, T0);
  CHECK(err.has_value());  // processHealthCheck should return an error for an invalid cron expression
  CHECK(err && err->find("expected exactly 5 fields") != std::string::npos);
}

static void TestProcessHealthCheck_ValidCron_SetsRepeatAfterSec(HealthCheckReconciler& r) {
  HealthCheck hc = newHC("valid-cron", "default");
  hc.Spec.RepeatAfterSec = 0;
  hc.Spec.Schedule.Cron = "@every 5s";
  auto err = r.ProcessHealthCheck(&hc, T0);
  CHECK(!err.has_value());
  CHECK(hc.Spec.RepeatAfterSec > 0);   // the reference's assertion
  CHECK(hc.Spec.RepeatAfterSec == 5);  // what hcc.go:259-262 intends
  HealthCheck five = newHC("five-field", "default");
  five.Spec.Schedule.Cron = "*/15 * * * *";
  CHECK(!r.ProcessHealthCheck(&five, T0).has_value());
  CHECK(five.Spec.RepeatAfterSec == 900);  // 09:15:00 -> 09:30:00
}

static void TestPausedHealthCheck_StatusStopped(HealthCheckReconciler& r) {
  HealthCheck hc = newHC("inline-hello-pause", "health");  // examples/bdd/inlineHelloTest.yaml: repeatAfterSec: 0
  hc.Spec.RepeatAfterSec = 0;
  CHECK(!r.Reconcile(hc.Key(), &hc, T0).has_value());
  bool stopped = false;
  for (const Due& d : r.Tick(T0)) stopped |= d.Key == hc.Key() && (d.Action & AM_ACT_STOPPED);
  CHECK(stopped);
  auto st = r.GetStatus(hc.Key());
  CHECK(st && st->Status == "Stopped" && st->FinishedAt && *st->FinishedAt == T0);
  for (const Due& d : r.Tick(T0 + 1)) CHECK(d.Key != hc.Key());  // reported once
  CHECK(!r.GetTimerByName(hc.Key()).has_value());
}

static void TestNilWorkflowResource_StatusUntouched(HealthCheckReconciler& r) {
  HealthCheck hc = newHC("nil-resource", "health");
  hc.Spec.RepeatAfterSec = 60;
  hc.Spec.Workflow.HasResource = false;
  CHECK(!r.Reconcile(hc.Key(), &hc, T0).has_value());
  for (const Due& d : r.Tick(T0 + 2)) CHECK(d.Key != hc.Key());
  auto st = r.GetStatus(hc.Key());
  CHECK(st && st->Status.empty() && !st->FinishedAt && st->TotalHealthCheckRuns == 0);
}

static void TestRemedyFlow_RunsLimitAndResetInterval(HealthCheckReconciler& r) {
  // examples/bdd/inlineMemoryRemedyUnitTest.yaml: repeatAfterSec 5, remedyRunsLimit 2,
  // remedyResetInterval 300; no Argo controller runs in envtest, so every workflow "fails".
  HealthCheck hc = newHC("inline-memory-remedy", "health");
  hc.Spec.RepeatAfterSec = 5;
  hc.Spec.Remedy.GenerateName = "inline-memory-remedy-";
  hc.Spec.Remedy.HasResource = true;
  hc.Spec.RemedyRunsLimit = 2;
  hc.Spec.RemedyResetInterval = 300;
  CHECK(!hc.Spec.Remedy.IsEmpty());
  CHECK(!r.Reconcile(hc.Key(), &hc, T0).has_value());
  int submits = 0, remedies = 0, skips = 0;
  for (int64_t t = T0 + 10; t < T0 + 70; ++t) {
    for (const Due& d : r.Tick(t)) {
      if (d.Key != hc.Key()) continue;
      if (d.Action & AM_ACT_SUBMIT_HC) { ++submits; r.PostResult(hc.Key(), Failed, Failed); }
      remedies += (d.Action & AM_ACT_RUN_REMEDY) != 0;
      skips += (d.Action & AM_ACT_REMEDY_SKIP) != 0;
    }
  }
  auto st = r.GetStatus(hc.Key());
  CHECK(st.has_value());
  CHECK(st && st->SuccessCount + st->FailedCount >= 3);  // the reference's liveness bound (test.go:68)
  CHECK(submits >= 9 && submits <= 12);                  // a 5 s check over 60 s (+1 s per result hop)
  CHECK(remedies == 2);                                  // remedyRunsLimit
  CHECK(skips == submits - 2 - 1 || skips == submits - 2);  // every later failure inside the reset interval
  CHECK(st && st->RemedyTotalRuns == 2 && st->RemedyFailedCount == 2);
  auto timer = r.GetTimerByName(hc.Key());
  CHECK(timer && st && st->FinishedAt && *timer == *st->FinishedAt + 5);  // hcc.go:751
}

static void TestDeleteStopsTimer(HealthCheckReconciler& r) {
  HealthCheck hc = newHC("to-be-deleted", "health");
  hc.Spec.RepeatAfterSec = 60;
  hc.Status.FinishedAt = T0;
  CHECK(!r.Reconcile(hc.Key(), &hc, T0).has_value());
  r.Tick(T0 + 100);                                  // submitted (elapsed >= 60, and no timer yet: hcc.go:264)
  CHECK(!r.GetTimerByName(hc.Key()).has_value());
  r.PostResult(hc.Key(), amsweep::Succeeded);
  r.Tick(T0 + 101);                                  // the result arms the repeat timer (hcc.go:745-752)
  CHECK(r.GetTimerByName(hc.Key()).has_value());
  CHECK(!r.Reconcile(hc.Key(), nullptr, T0 + 102).has_value());  // CR not found (hcc.go:175-186)
  CHECK(!r.GetTimerByName(hc.Key()).has_value());
  for (int64_t t = T0 + 103; t < T0 + 300; t += 60)
    for (const Due& d : r.Tick(t)) CHECK(d.Key != hc.Key());
}

int main() {
  std::string err;
  auto r = HealthCheckReconciler::New(0, 4096, &err);
  if (!r) { std::printf("no reconciler: %s\n", err.c_str()); return 2; }
  TestProcessHealthCheck_InvalidCron_ReturnsError(*r);
  TestProcessHealthCheck_ValidCron_SetsRepeatAfterSec(*r);
  TestPausedHealthCheck_StatusStopped(*r);
  TestNilWorkflowResource_StatusUntouched(*r);
  TestRemedyFlow_RunsLimitAndResetInterval(*r);
  TestDeleteStopsTimer(*r);
  std::printf(failures ? "FAILED %d checks\n" : "ok\n", failures);
  return failures ? 1 : 0;
}
