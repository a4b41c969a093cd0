// ingest_json.cpp — HealthCheck manifests (JSON, as the API server serves them) -> packed records,
// natively and on every host thread (the step BEFORE the path).
//
// At controller start the informer's initial list replays every CR through Reconcile (hcc.go:170);
// with millions of checks the field extraction must not be a per-record interpreter loop (round 1:
// Python dict walks, minutes for 10 M CRs).  One pass finds the spans of the HealthCheck objects
// in the document — a single object, a JSON array, or a `List` with "items" — then the spans are
// parsed in parallel: a small recursive-descent walker that only looks at the fields the ladder
// needs, by the JSON tags of api/v1alpha1/healthcheck_types.go:32-66, :88-102:
//
//   spec.repeatAfterSec  spec.schedule.cron  spec.workflow.resource (nil-ness, hcc.go:227)
//   spec.remedyworkflow.{generateName, resource, workflowtimeout, rbacRules}  (IsEmpty, :104-106)
//   spec.remedyRunsLimit  spec.remedyResetInterval
//   status.{finishedAt, remedyFinishedAt, successCount, failedCount, remedySuccessCount,
//           remedyFailedCount, remedyTotalRuns}
//
// and hands the result to am_healthcheck_classify (the ladder itself, cron_parse.cpp).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/amsweep.h"
#include "civil.h"

namespace {

struct Cur {
  const char* p;
  const char* e;
  bool ok = true;
  void ws() { while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  char peek() { ws(); return p < e ? *p : 0; }
};

void skip_value(Cur& c);

// after the opening quote; leaves p after the closing quote.  out (may be null) receives the decoded text.
void read_string(Cur& c, std::string* out) {
  while (c.p < c.e) {
    const char ch = *c.p++;
    if (ch == '"') return;
    if (ch != '\\') { if (out) out->push_back(ch); continue; }
    if (c.p >= c.e) break;
    const char esc = *c.p++;
    uint32_t cp = 0;
    switch (esc) {
      case '"': cp = '"'; break;   case '\\': cp = '\\'; break;  case '/': cp = '/'; break;
      case 'b': cp = '\b'; break;  case 'f': cp = '\f'; break;   case 'n': cp = '\n'; break;
      case 'r': cp = '\r'; break;  case 't': cp = '\t'; break;
      case 'u': {
        auto hex4 = [&](uint32_t& v) {
          v = 0;
          for (int k = 0; k < 4; ++k) {
            if (c.p >= c.e) return false;
            const char h = *c.p++;
            v <<= 4;
            if (h >= '0' && h <= '9') v |= (uint32_t)(h - '0');
            else if (h >= 'a' && h <= 'f') v |= (uint32_t)(h - 'a' + 10);
            else if (h >= 'A' && h <= 'F') v |= (uint32_t)(h - 'A' + 10);
            else return false;
          }
          return true;
        };
        if (!hex4(cp)) { c.ok = false; return; }
        if (cp >= 0xD800 && cp <= 0xDBFF && c.p + 6 <= c.e && c.p[0] == '\\' && c.p[1] == 'u') {
          const char* save = c.p;
          c.p += 2;
          uint32_t lo = 0;
          if (hex4(lo) && lo >= 0xDC00 && lo <= 0xDFFF) cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          else c.p = save;
        }
        break;
      }
      default: c.ok = false; return;
    }
    if (!out) continue;
    if (cp < 0x80) out->push_back((char)cp);
    else if (cp < 0x800) { out->push_back((char)(0xC0 | (cp >> 6))); out->push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
      out->push_back((char)(0xE0 | (cp >> 12))); out->push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out->push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      out->push_back((char)(0xF0 | (cp >> 18))); out->push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out->push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out->push_back((char)(0x80 | (cp & 0x3F)));
    }
  }
  c.ok = false;  // unterminated
}

void skip_container(Cur& c, char open, char close) {  // after `open`
  int depth = 1;
  while (c.p < c.e && depth) {
    const char ch = *c.p++;
    if (ch == '"') read_string(c, nullptr);
    else if (ch == open) ++depth;
    else if (ch == close) --depth;
  }
  if (depth) c.ok = false;
}

void skip_value(Cur& c) {
  c.ws();
  if (c.p >= c.e) { c.ok = false; return; }
  const char ch = *c.p;
  if (ch == '"') { ++c.p; read_string(c, nullptr); }
  else if (ch == '{') { ++c.p; skip_container(c, '{', '}'); }
  else if (ch == '[') { ++c.p; skip_container(c, '[', ']'); }
  else while (c.p < c.e && *c.p != ',' && *c.p != '}' && *c.p != ']' && *c.p != ' ' && *c.p != '\n' && *c.p != '\r' && *c.p != '\t') ++c.p;
}

// value kinds the extractor cares about
bool is_null(Cur& c) { c.ws(); return c.e - c.p >= 4 && memcmp(c.p, "null", 4) == 0; }

bool read_int(Cur& c, int64_t& v) {  // JSON number (integers; a fraction / exponent is truncated towards zero like int(float))
  c.ws();
  const char* s = c.p;
  skip_value(c);
  if (s == c.p) return false;
  if (*s == '"' || *s == '{' || *s == '[' || *s == 't' || *s == 'f' || *s == 'n') { v = 0; return *s == 'n'; }
  bool neg = false;
  const char* q = s;
  if (*q == '-') { neg = true; ++q; }
  unsigned long long acc = 0;
  bool any = false, over = false;
  while (q < c.p && *q >= '0' && *q <= '9') {
    any = true;
    if (acc > (0x7FFFFFFFFFFFFFFFull - 9) / 10) over = true;
    acc = acc * 10 + (unsigned long long)(*q - '0');
    ++q;
  }
  if (!any) return false;
  if (over) acc = 0x7FFFFFFFFFFFFFFFull;
  v = neg ? -(int64_t)acc : (int64_t)acc;
  return true;
}

// metav1.Time: RFC 3339, e.g. 2026-09-21T09:15:00Z / 2026-09-21T11:15:00.25+02:00 -> unix seconds (floor)
bool rfc3339(const std::string& s, int64_t& out) {
  auto num = [&](size_t pos, size_t n, int& v) {
    if (pos + n > s.size()) return false;
    v = 0;
    for (size_t k = 0; k < n; ++k) { if (s[pos + k] < '0' || s[pos + k] > '9') return false; v = v * 10 + (s[pos + k] - '0'); }
    return true;
  };
  int Y, M, D, h, m, sec;
  if (s.size() < 20 || !num(0, 4, Y) || s[4] != '-' || !num(5, 2, M) || s[7] != '-' || !num(8, 2, D) ||
      (s[10] != 'T' && s[10] != 't' && s[10] != ' ') || !num(11, 2, h) || s[13] != ':' || !num(14, 2, m) || s[16] != ':' || !num(17, 2, sec))
    return false;
  size_t pos = 19;
  if (pos < s.size() && s[pos] == '.') { ++pos; while (pos < s.size() && s[pos] >= '0' && s[pos] <= '9') ++pos; }
  int off = 0;
  if (pos < s.size() && (s[pos] == 'Z' || s[pos] == 'z')) ++pos;
  else if (pos < s.size() && (s[pos] == '+' || s[pos] == '-')) {
    int oh, om;
    if (!num(pos + 1, 2, oh) || pos + 3 >= s.size() || s[pos + 3] != ':' || !num(pos + 4, 2, om)) return false;
    off = (s[pos] == '-' ? -1 : 1) * (oh * 3600 + om * 60);
    pos += 6;
  } else return false;
  if (pos != s.size() || M < 1 || M > 12 || D < 1 || D > 31) return false;
  out = amsweep::days_from_civil(Y, M, D) * 86400 + h * 3600 + m * 60 + sec - off;
  return true;
}

// walk the members of an object: fn(key, cursor positioned at the value) must consume the value
template <class F>
void for_members(Cur& c, F fn) {
  if (!c.eat('{')) { skip_value(c); return; }
  if (c.eat('}')) return;
  std::string key;
  while (c.ok) {
    if (!c.eat('"')) { c.ok = false; return; }
    key.clear();
    read_string(c, &key);
    if (!c.eat(':')) { c.ok = false; return; }
    fn(key, c);
    if (c.eat(',')) continue;
    if (c.eat('}')) return;
    c.ok = false;
    return;
  }
}

struct Fields {
  int64_t ras = 0, limit = 0, reset = 0, s = 0, f = 0, rs = 0, rf = 0, rt = 0, timeout = 0;
  std::string cron, gen, fin, rfin;
  bool has_resource = false, rem_resource = false, rem_rbac = false, fin_set = false, rfin_set = false;
};

void parse_healthcheck(Cur& c, Fields& x) {
  for_members(c, [&](const std::string& k, Cur& c) {
    if (k == "spec") {
      for_members(c, [&](const std::string& k, Cur& c) {
        if (k == "repeatAfterSec") { if (!read_int(c, x.ras)) c.ok = false; }
        else if (k == "remedyRunsLimit") { if (!read_int(c, x.limit)) c.ok = false; }
        else if (k == "remedyResetInterval") { if (!read_int(c, x.reset)) c.ok = false; }
        else if (k == "schedule") {
          for_members(c, [&](const std::string& k, Cur& c) {
            if (k == "cron" && c.peek() == '"') { ++c.p; x.cron.clear(); read_string(c, &x.cron); }
            else skip_value(c);
          });
        } else if (k == "workflow") {
          for_members(c, [&](const std::string& k, Cur& c) {
            if (k == "resource") x.has_resource = !is_null(c);  // hcc.go:227
            skip_value(c);
          });
        } else if (k == "remedyworkflow") {
          for_members(c, [&](const std::string& k, Cur& c) {
            if (k == "generateName" && c.peek() == '"') { ++c.p; x.gen.clear(); read_string(c, &x.gen); return; }
            if (k == "workflowtimeout") { if (!read_int(c, x.timeout)) c.ok = false; return; }
            if (k == "resource") x.rem_resource = !is_null(c);
            if (k == "rbacRules") x.rem_rbac = !is_null(c);
            skip_value(c);
          });
        } else skip_value(c);
      });
    } else if (k == "status") {
      for_members(c, [&](const std::string& k, Cur& c) {
        auto time_field = [&](std::string& dst, bool& set) {
          if (c.peek() == '"') { ++c.p; dst.clear(); read_string(c, &dst); set = true; }
          else skip_value(c);
        };
        if (k == "finishedAt") time_field(x.fin, x.fin_set);
        else if (k == "remedyFinishedAt") time_field(x.rfin, x.rfin_set);
        else if (k == "successCount") { if (!read_int(c, x.s)) c.ok = false; }
        else if (k == "failedCount") { if (!read_int(c, x.f)) c.ok = false; }
        else if (k == "remedySuccessCount") { if (!read_int(c, x.rs)) c.ok = false; }
        else if (k == "remedyFailedCount") { if (!read_int(c, x.rf)) c.ok = false; }
        else if (k == "remedyTotalRuns") { if (!read_int(c, x.rt)) c.ok = false; }
        else skip_value(c);
      });
    } else skip_value(c);
  });
}

int ingest_one(const char* p, const char* e, uint32_t timer_armed, am_record_t* out) {
  Cur c{p, e};
  Fields x;
  parse_healthcheck(c, x);
  if (!c.ok) return AM_E_PARSE;
  am_healthcheck_t hc;
  memset(&hc, 0, sizeof hc);
  hc.repeat_after_sec = x.ras;
  hc.cron = x.cron.data();
  hc.cron_len = x.cron.size();
  hc.has_resource = x.has_resource;
  hc.has_remedy = !am_remedy_is_empty(x.gen.size(), !x.rem_resource, x.timeout, !x.rem_rbac);
  hc.remedy_runs_limit = x.limit;
  hc.remedy_reset_interval = x.reset;
  if (x.fin_set) { if (!rfc3339(x.fin, hc.finished_at)) return AM_E_PARSE; hc.finished_at_set = 1; }
  if (x.rfin_set) { if (!rfc3339(x.rfin, hc.remedy_finished_at)) return AM_E_PARSE; hc.remedy_finished_at_set = 1; }
  hc.success_count = x.s; hc.failed_count = x.f;
  hc.remedy_success_count = x.rs; hc.remedy_failed_count = x.rf; hc.remedy_total_runs = x.rt;
  hc.timer_armed = timer_armed;
  return am_healthcheck_classify(&hc, out);
}

}  // namespace

extern "C" int am_healthcheck_ingest_json(const char* json, size_t len, uint32_t timer_armed, am_record_t* out,
                                          int32_t* rc_out, uint64_t cap, uint64_t* n_out, int n_threads) {
  if (!json || !n_out || (cap && !out)) return AM_E_INVAL;
  *n_out = 0;
  // 1. spans of the HealthCheck objects: the document itself, the elements of a top-level array, or
  //    the elements of the "items" array of a List object.  One flat pass over the bytes (depth, in-string
  //    and escape state only): this is the serial part, everything else runs on all threads.
  std::vector<std::pair<const char*, const char*>> spans;
  {
    const char* p = json;
    const char* const e = json + len;
    while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r')) ++p;
    if (p >= e || (*p != '[' && *p != '{')) return AM_E_PARSE;
    const bool top_array = *p == '[';
    int depth = 0;            // counts both {} and []
    int items_depth = top_array ? 1 : -1;  // depth INSIDE the array that holds the HealthChecks
    const char* item_start = nullptr;
    const char* key_start = nullptr;       // last string token seen at depth 1 of a top-level object
    size_t key_len = 0;
    bool expect_items_array = false;
    const char* doc_start = p;
    for (; p < e; ++p) {
      const char ch = *p;
      if (ch == '"') {  // skip the string
        const char* s = p + 1;
        ++p;
        while (p < e && *p != '"') { if (*p == '\\') ++p; ++p; }
        if (p >= e) return AM_E_PARSE;
        if (!top_array && depth == 1 && items_depth < 0) { key_start = s; key_len = (size_t)(p - s); }
        continue;
      }
      if (ch == ':' && !top_array && depth == 1 && items_depth < 0) {
        expect_items_array = key_len == 5 && memcmp(key_start, "items", 5) == 0;
        continue;
      }
      if (ch == '{' || ch == '[') {
        if (expect_items_array && depth == 1) {
          if (ch == '[') items_depth = 2;
          expect_items_array = false;
        }
        if (ch == '{' && depth == items_depth) item_start = p;
        ++depth;
      } else if (ch == '}' || ch == ']') {
        --depth;
        if (depth < 0) return AM_E_PARSE;
        if (ch == '}' && depth == items_depth && item_start) { spans.emplace_back(item_start, p + 1); item_start = nullptr; }
        if (depth == 0) { ++p; break; }
      } else if (ch != ' ' && ch != '\n' && ch != '\t' && ch != '\r' && ch != ',') {
        expect_items_array = false;  // a scalar value after "items": not a List
      }
    }
    if (depth != 0) return AM_E_PARSE;
    if (!top_array && items_depth < 0) spans.emplace_back(doc_start, p);  // a single HealthCheck object
  }
  const uint64_t n = spans.size();
  *n_out = n;
  if (n > cap) return AM_E_NOSPACE;
  // 2. parse + classify, spans split over the host threads
  unsigned nt = n_threads > 0 ? (unsigned)n_threads : std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  if (nt > n) nt = n ? (unsigned)n : 1;
  auto work = [&](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) {
      const int rc = ingest_one(spans[i].first, spans[i].second, timer_armed, &out[i]);
      if (rc_out) rc_out[i] = rc;
    }
  };
  if (nt <= 1) { work(0, n); return AM_OK; }
  std::vector<std::thread> th;
  const uint64_t chunk = (n + nt - 1) / nt;
  for (unsigned t = 0; t < nt; ++t) {
    const uint64_t lo = (uint64_t)t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    if (lo >= hi) break;
    th.emplace_back(work, lo, hi);
  }
  for (auto& t : th) t.join();
  return AM_OK;
}
