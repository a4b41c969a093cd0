/*
 * amsweep_oracle.h — CPU ORACLE for the schedule-evaluation sweep.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference leg may build, load or call it, and only as the checker
 * or the reported CPU baseline.  libamsweep never links or calls it.
 *
 * What it restates (all file:line relative to /root/reference):
 *   - internal/controllers/healthcheck_controller.go ("hcc.go")
 *       :225-267  schedule ladder            -> orc_classify, orc_tick_record
 *       :607-611, :745-752 interval capture + re-arm -> due predicate
 *       :633-724  result + remedy gate       -> orc_tick_record step 1
 *       :819-852  remedy result              -> orc_tick_record step 1
 *   - api/v1alpha1/healthcheck_types.go :32-66, :104-106 (fields, IsEmpty)
 *   - github.com/robfig/cron/v3 v3.0.1 (go.mod:14; NOT in /root/reference):
 *       parser.go (Parse, getField, getRange, getBits, parseDescriptor),
 *       spec.go (SpecSchedule.Next, dayMatches), constantdelay.go (Every),
 *       restated from the published source as recorded in SURVEY.md App. A,
 *       plus Go's time.ParseDuration and strconv.Atoi acceptance rules.
 *
 * PARITY UNPINNED for 5-field cron expressions and for the numeric edges of
 * the remedy gate: no Go toolchain exists in this environment, robfig/cron is
 * not vendored, and the reference's own tests pin only (a) "NOT_A_VALID_CRON"
 * is rejected, (b) "@every 5s" yields RepeatAfterSec > 0, (c) the pause rule,
 * (d) RemedyWorkflow.IsEmpty's truth table.  Those four are
 * checked in tests/test_oracle_golden.py; everything else is pinned by the
 * independent Python restatement (oracle/oracle_py.py) and algebraic
 * properties, not by the reference.
 *
 * Calendar arithmetic deliberately uses libc gmtime_r/timegm so that it shares
 * no code with the product's closed-form civil-time routine.
 */
#ifndef AMSWEEP_ORACLE_H_
#define AMSWEEP_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* layout-identical to am_cron_t / am_healthcheck_t / am_record_t /
 * am_record_cols_t / am_tick_stats_t in include/amsweep.h (tests assert the
 * sizes); redeclared here so the oracle compiles without the product header */
typedef struct orc_cron {
  uint64_t minute, hour, dom, month, dow;
  int64_t delay_sec;
  int32_t kind;
  int32_t tz_id;
} orc_cron_t;

typedef struct orc_healthcheck {
  int64_t repeat_after_sec;
  const char* cron;
  size_t cron_len;
  int32_t has_resource;
  int32_t has_remedy;
  int64_t remedy_runs_limit;
  int64_t remedy_reset_interval;
  int64_t finished_at;
  int64_t remedy_finished_at;
  int32_t finished_at_set;
  int32_t remedy_finished_at_set;
  int64_t success_count, failed_count;
  int64_t remedy_success_count, remedy_failed_count, remedy_total_runs;
  uint32_t fail_p8;
  uint32_t timer_armed; /* r.GetTimerByName(name) != nil (hcc.go:264) */
} orc_healthcheck_t;

typedef struct orc_record {
  uint64_t minute, hour, dom, month, dow;
  int64_t finished_at;
  int64_t remedy_finished_at;
  int32_t ras;
  uint32_t flags;
  int32_t runs_limit, reset_interval;
  int32_t success, failed, remedy_success, remedy_failed, remedy_total;
  int32_t reserved;
} orc_record_t;

typedef struct orc_record_cols {
  uint64_t *minute, *hour, *dom, *month, *dow;
  int32_t* ras;
  uint32_t* flags;
  int64_t* finished_at;
  int32_t *runs_limit, *reset_interval;
  int32_t *success, *failed, *remedy_success, *remedy_failed, *remedy_total;
  int64_t* remedy_finished_at;
} orc_record_cols_t;

typedef struct orc_tick_stats {
  uint64_t n_records, n_emitted;
  uint64_t n_submit_hc, n_run_remedy, n_stopped, n_parse_error;
  uint64_t n_remedy_skip, n_reset_on_pass, n_reset_on_interval, n_anomaly;
  uint64_t n_result_ok, n_result_fail, n_remedy_ok, n_remedy_fail;
  uint64_t idx_xor, idx_sum;
} orc_tick_stats_t;

/* robfig/cron v3.0.1 ParseStandard.  0 ok, -6 rejected (message in err). */
int orc_cron_parse(const char* spec, size_t len, orc_cron_t* out, char* err, size_t errcap);
/* time.LoadLocation for "TZ=" / "CRON_TZ=" prefixes: ids in order of first appearance */
int orc_tz_lookup(const char* name, size_t len, int32_t* id);
int orc_tz_offset(int32_t id, int64_t unix_sec, int32_t* utoff); /* through libc's tz database */
int orc_cron_matches(const orc_cron_t* c, int64_t unix_sec);
int64_t orc_cron_next(const orc_cron_t* c, int64_t unix_sec);
int64_t orc_cron_repeat_after_sec(const orc_cron_t* c, int64_t unix_sec);
/* time.ParseDuration: 0 ok (*ns_out set), -1 error */
int orc_parse_duration(const char* s, size_t len, int64_t* ns_out);

int orc_classify(const orc_healthcheck_t* hc, orc_record_t* out);
int orc_remedy_is_empty(size_t generate_name_len, int resource_is_nil, int64_t timeout,
                        int rbac_rules_is_nil);

/* SURVEY Appendix B.3 for ONE record at tick T; returns the action byte. */
uint32_t orc_tick_record(orc_record_t* r, int64_t T, uint32_t mode, uint64_t seed,
                         uint64_t global_idx, orc_tick_stats_t* stats);

/* Whole-array tick, single thread, mutates cols in place.  Emits ascending
 * (global idx, action) pairs; returns 0 or -3 when cap is too small. */
int orc_sweep(orc_record_cols_t* cols, uint64_t n, uint64_t shard_base, int64_t T, uint32_t mode,
              uint64_t seed, uint64_t* due_idx, uint32_t* due_action, uint64_t cap,
              uint64_t* n_out, orc_tick_stats_t* stats);

/* Same, split over nthreads contiguous chunks (CPU baseline B2/B3). */
int orc_sweep_mt(orc_record_cols_t* cols, uint64_t n, uint64_t shard_base, int64_t T,
                 uint32_t mode, uint64_t seed, uint64_t* due_idx, uint32_t* due_action,
                 uint64_t cap, uint64_t* n_out, orc_tick_stats_t* stats, int nthreads);

/* CPU baseline B1 ("faithful shape", hcc.go:253-264): for each HealthCheck
 * re-parse the cron string, call Next(), run the ladder; returns #submits. */
uint64_t orc_faithful_eval(const orc_healthcheck_t* hcs, uint64_t n, int64_t T);

/* keyed splitmix64 shared by generator, oracle and kernel */
uint64_t orc_key(uint64_t seed, uint64_t i, uint64_t f);
void orc_civil_from_unix(int64_t unix_sec, int32_t out[6]);

#ifdef __cplusplus
}
#endif
#endif
