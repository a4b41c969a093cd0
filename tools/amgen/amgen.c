/*
 * amgen.c — deterministic synthetic HealthCheck populations.
 *
 * Neutral test/bench tooling: it knows nothing about cron semantics.  It draws
 * HealthCheck specs (repeatAfterSec / schedule.cron STRINGS / remedy knobs /
 * status) shaped like the reference's examples (examples/inlineHello.yaml:7-16,
 * examples/bdd/inlineMemoryRemedyUnitTest.yaml:8-11) from a keyed splitmix64
 * stream, and hands each one to a caller-supplied classify function — the
 * product's am_healthcheck_classify or the oracle's orc_classify, which share
 * the am_healthcheck_t / am_record_t layouts — to obtain the packed record.
 *
 * Populations (config id):
 *    1  N x {repeatAfterSec: 60}                         BASELINE configs[0]
 *   11  N x {cron: "@every 1m"} (inlineHello.yaml as shipped) — same due-set
 *    2  mixed 5-field cron + repeatAfterSec               BASELINE configs[1]
 *    3  config 2 + remedy state, 50 % pending Failed      BASELINE configs[2]
 *    4  = 2 (sharded by the caller)                       BASELINE configs[3]
 *    5  = 2 with per-record failure probability (closed loop), 55 = 3-mix
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdatomic.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "../../include/amsweep.h"

typedef int (*amgen_classify_fn)(const am_healthcheck_t*, am_record_t*);

static inline uint64_t sm64(uint64_t z) {
  z += 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
uint64_t amgen_key(uint64_t seed, uint64_t i, uint64_t f) { return sm64(sm64(seed ^ sm64(i)) + f); }

/* field ids of the keyed stream */
enum { K_KIND = 1, K_RAS, K_FIN, K_FINSET, K_FAILP, K_CRON = 16, K_REMEDY = 64, K_LIMIT, K_RESET,
       K_PEND, K_RT, K_RS, K_RFA, K_ROUT, K_S, K_F, K_VIOL, K_ARMED };

typedef struct { uint64_t seed, i, ctr; } rng_t;
static uint64_t rnd(rng_t* r) { return amgen_key(r->seed, r->i, K_CRON + (r->ctr++)); }
static uint32_t below(rng_t* r, uint32_t n) { return (uint32_t)(rnd(r) % n); }

static const char* const MON[] = {"jan", "feb", "mar", "apr", "may", "jun",
                                  "jul", "aug", "sep", "oct", "nov", "dec"};
static const char* const DOW[] = {"sun", "mon", "tue", "wed", "thu", "fri", "sat"};

typedef struct { int lo, hi; const int* steps; int nsteps; const char* const* names; int name_base; int qmark; } fdesc_t;
static const int ST_MIN[] = {2, 3, 5, 10, 15, 20, 30};
static const int ST_HR[] = {2, 3, 4, 6, 8, 12};
static const int ST_DOM[] = {2, 5, 7, 10, 15};
static const int ST_MON[] = {2, 3, 4, 6};
static const int ST_DOW[] = {2, 3};
static const fdesc_t FD[5] = {
    {0, 59, ST_MIN, 7, NULL, 0, 0}, {0, 23, ST_HR, 6, NULL, 0, 0}, {1, 31, ST_DOM, 5, NULL, 0, 1},
    {1, 12, ST_MON, 4, MON, 1, 0},  {0, 6, ST_DOW, 2, DOW, 0, 1},
};

static char* put_val(char* p, const fdesc_t* d, int v, rng_t* r) {
  if (d->names && below(r, 10) == 0) {
    const char* nm = d->names[v - d->name_base];
    /* names are case-insensitive in robfig: vary the case */
    int up = (int)below(r, 3);
    for (int k = 0; k < 3; k++) *p++ = (up == 2 || (up == 1 && k == 0)) ? (char)(nm[k] - 32) : nm[k];
    return p;
  }
  return p + sprintf(p, "%d", v);
}

/* one cron field from the fixed grammar: '*' 40 %, '* /k' 20 %, single 15 %,
 * range 10 %, list 10 %, range/k 5 %; '?' replaces '*' for 5 % of dom/dow */
static char* put_field(char* p, const fdesc_t* d, rng_t* r) {
  uint32_t u = below(r, 100);
  int span = d->hi - d->lo + 1;
  if (u < 40) {
    *p++ = (d->qmark && below(r, 20) == 0) ? '?' : '*';
  } else if (u < 60) {
    p += sprintf(p, "*/%d", d->steps[below(r, (uint32_t)d->nsteps)]);
  } else if (u < 75) {
    p = put_val(p, d, d->lo + (int)below(r, (uint32_t)span), r);
  } else if (u < 85) {
    int a = d->lo + (int)below(r, (uint32_t)span);
    int b = a + (int)below(r, (uint32_t)(d->hi - a + 1));
    p = put_val(p, d, a, r);
    *p++ = '-';
    p = put_val(p, d, b, r);
  } else if (u < 95) {
    int cnt = 2 + (int)below(r, 3);
    for (int k = 0; k < cnt; k++) {
      if (k) *p++ = ',';
      p = put_val(p, d, d->lo + (int)below(r, (uint32_t)span), r);
    }
  } else {
    int a = d->lo + (int)below(r, (uint32_t)span);
    int b = a + (int)below(r, (uint32_t)(d->hi - a + 1));
    p += sprintf(p, "%d-%d/%d", a, b, d->steps[below(r, (uint32_t)d->nsteps)]);
  }
  return p;
}

static const char* const EVERY[] = {"@every 3s", "@every 5s", "@every 1m", "@every 90s",
                                    "@every 1h", "@every 1h30m", "@every 500ms"};
static const int64_t EVERY_SEC[] = {3, 5, 60, 90, 3600, 5400, 1};
static const char* const BAD[] = {"NOT_A_VALID_CRON", "0 * * * * *", "60 * * * *", "* * * * 7",
                                  "* * * *", "*/0 * * * *", "5-1 * * * *", "@every",
                                  "@fortnightly", "* 24 * * *", "* * 0 * *", "* * * 13 *",
                                  "1-2-3 * * * *", "1/2/3 * * * *", "a * * * *", "@every 5 s"};
static const char* const DESC[] = {"@hourly", "@daily", "@midnight", "@weekly", "@monthly",
                                   "@yearly", "@annually"};
static const int RAS_CHOICES[] = {5, 10, 30, 60, 300, 900, 3600};
static const char* const ZONES[] = {"America/New_York", "Europe/Paris", "Asia/Kolkata", "Asia/Kathmandu",
                                    "Australia/Lord_Howe", "America/St_Johns"};
static const char* const TZ_PREFIX[] = {"CRON_TZ=UTC ", "TZ=UTC ", "CRON_TZ=America/New_York ", "TZ=Europe/Paris ",
                                        "CRON_TZ=Asia/Kolkata ", "CRON_TZ=Asia/Kathmandu ",
                                        "TZ=Australia/Lord_Howe ", "CRON_TZ=America/St_Johns "};

#define AMGEN_STR 128 /* bytes reserved per cron string */

/* Draw HealthCheck #i of population `config`; cron text goes to buf[AMGEN_STR].
 * *post_flags receives pending-result bits to OR into the classified record
 * (config 3: the result this tick; not part of the CR). */
void amgen_healthcheck(int config, uint64_t seed, uint64_t i, int64_t T0, am_healthcheck_t* hc,
                       char* buf, uint32_t* post_flags) {
  memset(hc, 0, sizeof *hc);
  *post_flags = 0;
  buf[0] = 0;
  hc->has_resource = 1;
  hc->timer_armed = 1; /* configs 1 / 11: a running controller, every check has its repeat timer */
  hc->cron = buf;
  int64_t period = 60;
  int mix = (config == 3 || config == 55) ? 3 : 2;
  /* config 22 = config 2's population with 2 % of its 5-field specs bound to a time zone (the grammar
   * of SURVEY 8d config 2 has no prefix; the zone path has its own population) */
  const int zones = config == 22;
  if (config == 1) {
    hc->repeat_after_sec = 60;
    hc->finished_at_set = 1;
    hc->finished_at = T0 - (int64_t)(amgen_key(seed, i, K_FIN) % 120);
    return;
  }
  if (config == 11) {
    strcpy(buf, "@every 1m");
    hc->cron_len = strlen(buf);
    hc->finished_at_set = 1;
    hc->finished_at = T0 - (int64_t)(amgen_key(seed, i, K_FIN) % 120);
    return;
  }
  uint32_t u = (uint32_t)(amgen_key(seed, i, K_KIND) % 1000);
  rng_t r = {seed, i, 0};
  if (u < 500) { /* INTERVAL */
    hc->repeat_after_sec = RAS_CHOICES[amgen_key(seed, i, K_RAS) % 7];
    period = hc->repeat_after_sec;
    if (below(&r, 8) == 0) { /* cron present but ignored (hcc.go:251 vs :264) */
      strcpy(buf, "*/5 * * * *");
      hc->cron_len = strlen(buf);
    }
  } else if (u < 900) { /* 5-field cron / descriptor */
    period = 3600;
    if (below(&r, 25) == 0) {
      strcpy(buf, DESC[below(&r, 7)]);
    } else {
      char* p = buf;
      /* config 22: 2 % of the 5-field specs carry a time-zone prefix: UTC under both spellings and six named
       * zones, among them half-hour, 45-minute and 30-minute-DST offsets (robfig: time.LoadLocation) */
      if (below(&r, 50) == 0) {  /* (the draws are made for every config: configs 2 and 22 differ in the prefix only) */
        const uint64_t which = below(&r, 8);
        if (zones) p += sprintf(p, "%s", TZ_PREFIX[which]);
      }
      for (int f = 0; f < 5; f++) {
        if (f) { *p++ = ' '; if (below(&r, 40) == 0) *p++ = (below(&r, 2) ? '\t' : ' '); }
        p = put_field(p, &FD[f], &r);
      }
      *p = 0;
    }
    hc->cron_len = strlen(buf);
    hc->repeat_after_sec = -(int64_t)below(&r, 2); /* 0 or -1: both take the cron arm */
  } else if (u < 980) { /* @every */
    uint32_t k = below(&r, 7);
    strcpy(buf, EVERY[k]);
    hc->cron_len = strlen(buf);
    period = EVERY_SEC[k];
  } else if (u < 990) { /* paused */
    hc->repeat_after_sec = -(int64_t)below(&r, 2);
  } else if (u < 995) { /* parse error */
    strcpy(buf, BAD[below(&r, 16)]);
    hc->cron_len = strlen(buf);
  } else { /* Workflow.Resource == nil */
    hc->has_resource = 0;
    hc->repeat_after_sec = 60;
  }
  if (amgen_key(seed, i, K_FINSET) % 50 != 0) { /* 2 % never ran */
    hc->finished_at_set = 1;
    hc->finished_at = T0 - (int64_t)(amgen_key(seed, i, K_FIN) % (uint64_t)(2 * period + 1));
  }
  hc->fail_p8 = (uint32_t)(amgen_key(seed, i, K_FAILP) % 64);
  /* 2 % of the checks have no repeat timer (left over from a controller restart, hcc.go:161):
   * the ladder submits them whatever finishedAt says (hcc.go:264) */
  hc->timer_armed = (amgen_key(seed, i, K_ARMED) % 50) != 0;
  hc->success_count = (int64_t)(amgen_key(seed, i, K_S) % 1001);
  hc->failed_count = (int64_t)(amgen_key(seed, i, K_F) % 1001);

  if (mix == 3) {
    static const int LIM[] = {0, 1, 2, 5};
    static const int RST[] = {0, 60, 300};
    hc->has_remedy = (amgen_key(seed, i, K_REMEDY) % 100) < 60;
    hc->remedy_runs_limit = LIM[amgen_key(seed, i, K_LIMIT) % 4];
    hc->remedy_reset_interval = RST[amgen_key(seed, i, K_RESET) % 3];
    int64_t rt = (int64_t)(amgen_key(seed, i, K_RT) % 7);
    int64_t rs = (int64_t)(amgen_key(seed, i, K_RS) % (uint64_t)(rt + 1));
    hc->remedy_total_runs = rt;
    hc->remedy_success_count = rs;
    hc->remedy_failed_count = rt - rs;
    if (rt > 0) {
      hc->remedy_finished_at_set = 1;
      hc->remedy_finished_at = T0 - (int64_t)(amgen_key(seed, i, K_RFA) % 601);
    }
    if (amgen_key(seed, i, K_VIOL) % 1000 == 0) { /* N3: limit reached, nil time */
      hc->has_remedy = 1;
      hc->remedy_runs_limit = 2;
      hc->remedy_reset_interval = 300;
      hc->remedy_total_runs = 3;
      hc->remedy_success_count = 1;
      hc->remedy_failed_count = 2;
      hc->remedy_finished_at_set = 0;
      hc->remedy_finished_at = 0;
    }
    if (config == 3) {
      uint32_t pz = (uint32_t)(amgen_key(seed, i, K_PEND) % 4);
      if (pz < 2) {
        *post_flags |= AM_F_PENDING_FAIL | AM_F_REMEDY_PENDING;
        if ((amgen_key(seed, i, K_ROUT) % 10) < 7) *post_flags |= AM_F_REMEDY_OUTCOME_OK;
      } else if (pz == 2) {
        *post_flags |= AM_F_PENDING_OK;
      }
    }
  }
}

typedef struct {
  int config;
  uint64_t seed, first, lo, hi;
  int64_t T0;
  amgen_classify_fn fn;
  const am_record_cols_t* cols;
  int64_t bad; /* classify failures other than AM_E_UNSUPPORTED */
} job_t;

static void* fill_worker(void* p) {
  job_t* j = (job_t*)p;
  const am_record_cols_t* c = j->cols;
  char buf[AMGEN_STR];
  for (uint64_t i = j->lo; i < j->hi; i++) {
    am_healthcheck_t hc;
    am_record_t r;
    uint32_t post;
    amgen_healthcheck(j->config, j->seed, i, j->T0, &hc, buf, &post);
    int rc = j->fn(&hc, &r);
    if (rc != 0 && rc != AM_E_UNSUPPORTED) j->bad++;
    uint32_t kind = r.flags & AM_KIND_MASK;
    if (kind != AM_KIND_NO_RESOURCE && kind != AM_KIND_HOST_FALLBACK) r.flags |= post;
    uint64_t o = i - j->first;
    c->minute[o] = r.minute; c->hour[o] = r.hour; c->dom[o] = r.dom; c->month[o] = r.month;
    c->dow[o] = r.dow; c->ras[o] = r.ras; c->flags[o] = r.flags; c->finished_at[o] = r.finished_at;
    c->runs_limit[o] = r.runs_limit; c->reset_interval[o] = r.reset_interval;
    c->success[o] = r.success; c->failed[o] = r.failed; c->remedy_success[o] = r.remedy_success;
    c->remedy_failed[o] = r.remedy_failed; c->remedy_total[o] = r.remedy_total;
    c->remedy_finished_at[o] = r.remedy_finished_at;
  }
  return NULL;
}

/* Fill SoA columns for records [first, first+n) of a population through the
 * given classify function.  Returns the number of unexpected classify
 * failures (0 expected). */
int64_t amgen_fill(int config, uint64_t seed, uint64_t first, uint64_t n, int64_t T0,
                   amgen_classify_fn fn, const am_record_cols_t* cols, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  /* Zone ids are handed out in order of first appearance by whichever implementation sits behind
   * `fn` (product or oracle): introduce the zones in a fixed order, single-threaded, before the
   * threaded fill, so that the flags columns of two implementations are comparable. */
  for (size_t z = 0; config == 22 && z < sizeof ZONES / sizeof ZONES[0]; z++) {
    am_healthcheck_t hc;
    am_record_t r;
    char spec[AMGEN_STR];
    memset(&hc, 0, sizeof hc);
    snprintf(spec, sizeof spec, "CRON_TZ=%s * * * * *", ZONES[z]);
    hc.has_resource = 1;
    hc.cron = spec;
    hc.cron_len = strlen(spec);
    (void)fn(&hc, &r);
  }
  job_t jobs[256];
  pthread_t th[256];
  uint64_t chunk = (n + (uint64_t)nthreads - 1) / (uint64_t)nthreads;
  for (int t = 0; t < nthreads; t++) {
    uint64_t lo = (uint64_t)t * chunk < n ? (uint64_t)t * chunk : n;
    uint64_t hi = lo + chunk < n ? lo + chunk : n;
    jobs[t] = (job_t){config, seed, first, first + lo, first + hi, T0, fn, cols, 0};
  }
  for (int t = 1; t < nthreads; t++) pthread_create(&th[t], NULL, fill_worker, &jobs[t]);
  fill_worker(&jobs[0]);
  int64_t bad = jobs[0].bad;
  for (int t = 1; t < nthreads; t++) { pthread_join(th[t], NULL); bad += jobs[t].bad; }
  return bad;
}

/* Materialise the HealthChecks themselves (for the Python oracle and for the
 * "faithful shape" CPU baseline): hcs[k].cron points into strpool + k*96. */
void amgen_healthchecks(int config, uint64_t seed, uint64_t first, uint64_t n, int64_t T0,
                        am_healthcheck_t* hcs, char* strpool, uint32_t* post_flags) {
  for (uint64_t k = 0; k < n; k++) {
    amgen_healthcheck(config, seed, first + k, T0, &hcs[k], strpool + k * AMGEN_STR,
                      &post_flags[k]);
  }
}

int amgen_str_stride(void) { return AMGEN_STR; }

/* Harness stand-in for the controller's loop over a tick's result (the Go shim
 * walks the (index, action) list to submit workflows, hcc.go:269-288): collect
 * the LOCAL slots of the checks whose action carries AM_ACT_SUBMIT_HC, i.e. the
 * workflows whose completion the closed-loop harness posts back next tick. */
uint64_t amgen_select_submitted(const uint64_t* idx, const uint32_t* act, uint64_t n, uint64_t base,
                                uint64_t* out_local) {
  uint64_t m = 0;
  for (uint64_t k = 0; k < n; k++) {
    out_local[m] = idx[k] - base;
    m += (act[k] & AM_ACT_SUBMIT_HC) ? 1u : 0u;
  }
  return m;
}

/* The same walk over an am_tick_view_t (u32 local indices, u8 actions in the library's
 * pinned buffer): slots of the submitted checks, as the u64 array am_sweep_post_result takes. */
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
uint64_t amgen_select_submitted_view(const uint32_t* idx_local, const uint8_t* act, uint64_t n,
                                     uint64_t* out_local) {
  uint64_t m = 0, k = 0;
  /* eight entries at a time: when all eight carry the bit (almost always: 98 % of the entries
   * are bare submits) the copy is a plain widening loop the compiler vectorises */
  for (; k + 8 <= n; k += 8) {
    uint64_t a8;
    memcpy(&a8, act + k, 8);
    if ((a8 & 0x0101010101010101ull) == 0x0101010101010101ull) {
      for (int j = 0; j < 8; j++) out_local[m + j] = idx_local[k + j];
      m += 8;
    } else {
      for (int j = 0; j < 8; j++) {
        out_local[m] = idx_local[k + j];
        m += (uint64_t)(act[k + j] & AM_ACT_SUBMIT_HC);
      }
    }
  }
  for (; k < n; k++) {
    out_local[m] = idx_local[k];
    m += (uint64_t)(act[k] & AM_ACT_SUBMIT_HC);
  }
  return m;
}

/* The host-closed loop of bench.py's e2e measurement, as the cgo shim would run it: compiled code
 * calling the C-ABI through function pointers (the harness does not link libamsweep).  Per step:
 * tick at the next second (am_sweep_tick_view: the GPU writes the list into the library's pinned
 * host buffer; on a multi-GPU shard am_gather_tick_view with `tick_handle` = the exchange: sweep,
 * NVLink exchange, this rank's own part of the global list); then the consumer side — `workers` threads, the controller's reconcile workers
 * (hcc.go:170-188, MaxConcurrentReconciles) — walks the list in pieces (hcc.go:269-288 stand-in),
 * each thread posting every submitted check of its piece as Succeeded (am_sweep_post_result, host
 * memory -> device) before the next tick.  Workers are persistent and spin between ticks.
 * Times `steps` steps after `warm` untimed ones with CLOCK_MONOTONIC; split[0..4) accumulates the
 * seconds in post (mean over workers) / tick / walk (mean over workers) / the whole consumer phase
 * (wall).  Returns 0 or the first failing return code. */
typedef int (*am_post_fn)(void*, uint64_t, const uint64_t*, const uint8_t*, const uint8_t*);
typedef int (*am_tick_view_fn)(void*, int64_t, uint32_t, am_tick_view_t*, am_tick_stats_t*);
static double mono_s(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
#define E2E_PIECE 16384u
#define E2E_MAX_WORKERS 16
typedef struct {
  am_post_fn post;
  void* handle;
  const uint8_t* ok_phase;
  uint64_t* slots; /* workers x piece scratch */
  uint64_t piece;          /* entries per piece of this tick's list */
  uint64_t scratch_stride; /* entries of scratch per worker */
  int workers;
  /* the tick's job */
  const uint32_t* idx;
  const uint8_t* act;
  uint64_t n;
  atomic_ullong next_piece, submitted, gen;
  atomic_int done, rc, quit;
  double post_s[E2E_MAX_WORKERS], walk_s[E2E_MAX_WORKERS];
} e2e_pool_t;
typedef struct { e2e_pool_t* pool; int w; } e2e_arg_t;

static void e2e_consume(e2e_pool_t* P, int w) {
  uint64_t* scratch = P->slots + (uint64_t)w * P->scratch_stride;
  uint64_t sub = 0;
  double walk = 0, post_s = 0;
  for (;;) {
    const uint64_t off = atomic_fetch_add_explicit(&P->next_piece, 1, memory_order_relaxed) * P->piece;
    if (off >= P->n) break;
    const uint64_t len = P->n - off < P->piece ? P->n - off : P->piece;
    const double w0 = mono_s();
    const uint64_t m = amgen_select_submitted_view(P->idx + off, P->act + off, len, scratch);
    const double w1 = mono_s();
    if (m) {
      const int rp = P->post(P->handle, m, scratch, P->ok_phase, NULL);
      if (rp) atomic_store(&P->rc, rp);
    }
    const double w2 = mono_s();
    walk += w1 - w0; post_s += w2 - w1;
    sub += m;
  }
  P->walk_s[w] += walk; P->post_s[w] += post_s;
  atomic_fetch_add_explicit(&P->submitted, sub, memory_order_relaxed);
}
/* CPUs for the consumer workers: one per physical core, on the calling thread's package (the pinned
 * buffers and the GPU's PCIe root are local to it), never the calling thread's own core — a worker
 * spinning on the ticking thread's SMT sibling slowed the tick by a third.  Returns how many were found
 * (0: topology not readable, the workers stay unpinned). */
static int e2e_topology(int cpu, const char* what) {
  char path[128];
  int v = -1;
  snprintf(path, sizeof path, "/sys/devices/system/cpu/cpu%d/topology/%s", cpu, what);
  FILE* f = fopen(path, "r");
  if (!f) return -1;
  if (fscanf(f, "%d", &v) != 1) v = -1;
  fclose(f);
  return v;
}
static int e2e_pick_cpus(int want, int* cpus, const cpu_set_t* allowed, int me) {
  const cpu_set_t set = *allowed;
  if (me < 0) return 0;
  const int my_pkg = e2e_topology(me, "physical_package_id"), my_core = e2e_topology(me, "core_id");
  if (my_pkg < 0 || my_core < 0) return 0;
  int seen[1024], n_seen = 0, n = 0;
  for (int cpu = 0; cpu < CPU_SETSIZE && n < want; cpu++) {
    if (!CPU_ISSET(cpu, &set)) continue;
    const int pkg = e2e_topology(cpu, "physical_package_id"), core = e2e_topology(cpu, "core_id");
    if (pkg != my_pkg || core < 0 || core == my_core) continue;
    int dup = 0;
    for (int k = 0; k < n_seen; k++) dup |= seen[k] == core;
    if (dup || n_seen >= 1024) continue;
    seen[n_seen++] = core;
    cpus[n++] = cpu;
  }
  return n;
}

static void* e2e_worker(void* arg) {
  e2e_pool_t* P = ((e2e_arg_t*)arg)->pool;
  const int w = ((e2e_arg_t*)arg)->w;
  unsigned long long seen = 0;
  for (;;) {
    unsigned long long g;
    while ((g = atomic_load_explicit(&P->gen, memory_order_acquire)) == seen) {
      if (atomic_load_explicit(&P->quit, memory_order_relaxed)) return NULL;
      /* a long pause between polls: a worker spinning hard on the ticking thread's SMT sibling (or its
       * cache line) cost the tick 36 us with ten workers; the price is <= ~1 us of wake-up latency */
      for (int q = 0; q < 64; q++) __builtin_ia32_pause();
    }
    seen = g;
    e2e_consume(P, w);
    atomic_fetch_add_explicit(&P->done, 1, memory_order_release);
  }
}

int amgen_e2e_closed_loop(am_post_fn post, am_tick_view_fn tick, void* handle, void* tick_handle, int64_t T_first, uint32_t mode,
                          uint64_t warm, uint64_t steps, uint64_t* slots /* capacity entries */, uint64_t capacity,
                          const uint8_t* ok_phase /* capacity x AM_PHASE_SUCCEEDED */, int workers, double* seconds,
                          double* split /* [4] */, uint64_t* h2d_bytes, uint64_t* d2h_bytes, uint64_t* last_emitted,
                          uint64_t* last_submitted) {
  uint64_t n_prev = 0, h2d = 0, d2h = 0;
  double t0 = 0, sp_tick = 0, sp_cons = 0;
  am_tick_view_t v;
  am_tick_stats_t st;
  memset(&v, 0, sizeof v);
  static e2e_pool_t P; /* (atomics: not copyable) */
  memset(&P, 0, sizeof P);
  P.post = post; P.handle = handle; P.ok_phase = ok_phase; P.slots = slots;
  /* pieces of 16 K list entries: small enough that the library copies finished pieces to the device while the
   * others are still walked (dividing the list evenly among the workers — one large post each — left the whole
   * copy to the tick: 0.25 -> 0.32 ms), large enough that the fixed cost of a post stays small */
  P.piece = E2E_PIECE;
  const char* pe = getenv("AMGEN_E2E_PIECE"); /* tests: force several pieces on a small population */
  if (pe && atoll(pe) > 0) P.piece = (uint64_t)atoll(pe);
  if (workers < 1) workers = 1;
  if (workers > E2E_MAX_WORKERS) workers = E2E_MAX_WORKERS;
  while (workers > 1 && (uint64_t)workers * P.piece > capacity) workers--;
  if (P.piece > capacity) P.piece = capacity;
  P.workers = workers;
  P.scratch_stride = capacity / (uint64_t)workers;  /* >= any piece: a piece never exceeds ceil(n / workers) <= this, or E2E_PIECE */
  pthread_t th[E2E_MAX_WORKERS];
  e2e_arg_t args[E2E_MAX_WORKERS];
  int cpus[E2E_MAX_WORKERS];
  cpu_set_t old_mask;
  int n_cpus = 0, repin = 0;
  /* opt-in (AMGEN_E2E_PIN=1): with ten workers pinning took the step from 0.78 to 0.48 ms on one box and to
   * 1.03 ms on the next (the package of the ticking thread is not always the one the pinned buffers live on) */
  if (getenv("AMGEN_E2E_PIN") && workers > 1 && sched_getaffinity(0, sizeof old_mask, &old_mask) == 0) {
    const int me = sched_getcpu();
    n_cpus = e2e_pick_cpus(workers - 1, cpus, &old_mask, me);
    if (n_cpus > 0) {  /* the ticking thread stays where it is */
      cpu_set_t mine;
      CPU_ZERO(&mine);
      CPU_SET(me, &mine);
      repin = sched_setaffinity(0, sizeof mine, &mine) == 0;
    }
  }
  for (int w = 1; w < workers; w++) {
    args[w].pool = &P; args[w].w = w;
    if (pthread_create(&th[w], NULL, e2e_worker, &args[w])) return -1;
    if (w - 1 < n_cpus) {
      cpu_set_t one;
      CPU_ZERO(&one);
      CPU_SET(cpus[w - 1], &one);
      pthread_setaffinity_np(th[w], sizeof one, &one);
    }
  }
  int rc = 0;
  for (uint64_t k = 0; k < warm + steps && !rc; k++) {
    if (k == warm) {
      t0 = mono_s(); h2d = d2h = 0; sp_tick = sp_cons = 0;
      for (int w = 0; w < workers; w++) P.post_s[w] = P.walk_s[w] = 0;
    }
    const double b = mono_s();
    rc = tick(tick_handle ? tick_handle : handle, T_first + (int64_t)k, mode, &v, &st);
    if (rc) break;
    const double c = mono_s();
    P.idx = v.idx_local; P.act = v.action; P.n = v.n;
    atomic_store(&P.next_piece, 0); atomic_store(&P.submitted, 0); atomic_store(&P.done, 0);
    atomic_fetch_add_explicit(&P.gen, 1, memory_order_release);
    e2e_consume(&P, 0);
    while (atomic_load_explicit(&P.done, memory_order_acquire) < workers - 1) __builtin_ia32_pause();
    rc = atomic_load(&P.rc);
    n_prev = atomic_load(&P.submitted);
    const double d = mono_s();
    h2d += n_prev * 8;
    d2h += v.n * 5 + sizeof st;
    sp_tick += c - b; sp_cons += d - c;
  }
  *seconds = mono_s() - t0;
  atomic_store(&P.quit, 1);
  for (int w = 1; w < workers; w++) pthread_join(th[w], NULL);
  if (repin) sched_setaffinity(0, sizeof old_mask, &old_mask);
  if (rc) return rc;
  if (split) {
    double ps = 0, ws = 0;
    for (int w = 0; w < workers; w++) { ps += P.post_s[w]; ws += P.walk_s[w]; }
    split[0] = ps / workers; split[1] = sp_tick; split[2] = ws / workers; split[3] = sp_cons;
  }
  *h2d_bytes = h2d; *d2h_bytes = d2h;
  *last_emitted = v.n; *last_submitted = n_prev;
  return workers;
}
