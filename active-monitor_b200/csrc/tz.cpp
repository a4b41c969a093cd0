// tz.cpp — TZif reader + POSIX TZ footer rules + the zone registry (see tz.h).
#include "tz.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

namespace amsweep_tz {
namespace {

using amsweep::civil_from_days;
using amsweep::days_from_civil;
using amsweep::split_days;
using amsweep::weekday_from_days;

struct Rule {      // one side of a POSIX TZ daylight rule
  int kind = -1;   // 0: Jn (1..365, Feb 29 never counted)  1: n (0..365)  2: Mm.w.d
  int a = 0, b = 0, c = 0;
  int32_t time = 7200;  // seconds after local midnight (default 02:00:00; may be negative or > 24 h)
};

struct Zone {
  std::string name;
  std::vector<int64_t> trans;  // transition instants, ascending
  std::vector<int32_t> off;    // off[i] is in force from trans[i] on
  int32_t off_first = 0;       // before the first transition
  bool has_footer = false, has_dst = false;
  int32_t std_off = 0, dst_off = 0;  // seconds EAST of UTC
  Rule start, end;
};

int64_t be64(const unsigned char* p) {
  uint64_t v = 0;
  for (int i = 0; i < 8; ++i) v = (v << 8) | p[i];
  return (int64_t)v;
}
int32_t be32(const unsigned char* p) {
  return (int32_t)(((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3]);
}

// ---- POSIX TZ string: std offset [dst [offset] [,start[/time],end[/time]]] ----
bool tz_name(const char*& s) {
  if (*s == '<') {
    ++s;
    while (*s && *s != '>') ++s;
    if (*s != '>') return false;
    ++s;
    return true;
  }
  int n = 0;
  while ((*s >= 'A' && *s <= 'Z') || (*s >= 'a' && *s <= 'z')) { ++s; ++n; }
  return n >= 3;
}
bool tz_hms(const char*& s, int32_t& out, bool allow_sign) {
  int sign = 1;
  if (allow_sign && (*s == '+' || *s == '-')) { if (*s == '-') sign = -1; ++s; }
  if (*s < '0' || *s > '9') return false;
  int32_t v[3] = {0, 0, 0};
  for (int k = 0; k < 3; ++k) {
    int n = 0;
    while (*s >= '0' && *s <= '9' && n < 4) { v[k] = v[k] * 10 + (*s - '0'); ++s; ++n; }
    if (n == 0) return false;
    if (*s != ':' || k == 2) break;
    ++s;
  }
  out = sign * (v[0] * 3600 + v[1] * 60 + v[2]);
  return true;
}
bool tz_rule(const char*& s, Rule& r) {
  if (*s == 'M') {
    ++s;
    r.kind = 2;
    int* f[3] = {&r.a, &r.b, &r.c};
    for (int k = 0; k < 3; ++k) {
      int v = 0, n = 0;
      while (*s >= '0' && *s <= '9') { v = v * 10 + (*s - '0'); ++s; ++n; }
      if (!n) return false;
      *f[k] = v;
      if (k < 2) { if (*s != '.') return false; ++s; }
    }
    if (r.a < 1 || r.a > 12 || r.b < 1 || r.b > 5 || r.c > 6) return false;
  } else {
    r.kind = 1;
    if (*s == 'J') { r.kind = 0; ++s; }
    int v = 0, n = 0;
    while (*s >= '0' && *s <= '9') { v = v * 10 + (*s - '0'); ++s; ++n; }
    if (!n) return false;
    r.a = v;
  }
  r.time = 7200;
  if (*s == '/') { ++s; if (!tz_hms(s, r.time, true)) return false; }
  return true;
}
bool parse_footer(const std::string& f, Zone& z) {
  const char* s = f.c_str();
  int32_t west = 0;
  if (!tz_name(s) || !tz_hms(s, west, true)) return false;
  z.std_off = -west;
  z.has_dst = false;
  if (*s == 0) return true;
  if (!tz_name(s)) return false;
  z.dst_off = z.std_off + 3600;
  if (*s && *s != ',') { if (!tz_hms(s, west, true)) return false; z.dst_off = -west; }
  if (*s == 0) return true;  // "std offset dst" without rules: POSIX leaves it to the implementation; treat as no DST
  if (*s != ',') return false;
  ++s;
  if (!tz_rule(s, z.start) || *s != ',') return false;
  ++s;
  if (!tz_rule(s, z.end) || *s != 0) return false;
  z.has_dst = true;
  return true;
}

bool is_leap(int64_t y) { return (y % 4 == 0 && y % 100 != 0) || y % 400 == 0; }

// seconds since the epoch of LOCAL midnight-based rule instant in year y (in the rule's own local clock)
int64_t rule_local_seconds(const Rule& r, int64_t y) {
  int64_t day;  // days since the epoch of the rule's date
  if (r.kind == 0) {
    int d = r.a;  // 1..365, Feb 29 never counted
    if (is_leap(y) && d >= 60) d += 1;
    day = days_from_civil(y, 1, 1) + d - 1;
  } else if (r.kind == 1) {
    day = days_from_civil(y, 1, 1) + r.a;
  } else {
    const int64_t first = days_from_civil(y, r.a, 1);
    const int wd = weekday_from_days(first);
    int d = 1 + (r.c - wd + 7) % 7 + 7 * (r.b - 1);
    static const int mdays[12] = {31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31};
    int len = mdays[r.a - 1] + ((r.a == 2 && is_leap(y)) ? 1 : 0);
    while (d > len) d -= 7;  // week 5 = the last one
    day = first + d - 1;
  }
  return day * 86400 + r.time;
}

int32_t footer_offset(const Zone& z, int64_t utc) {
  if (!z.has_dst) return z.std_off;
  int64_t days, y;
  int32_t sod, m, d;
  split_days(utc + z.std_off, days, sod);
  civil_from_days(days, y, m, d);
  // the rule instants of year y, as UTC: the start is given in standard time, the end in daylight time
  const int64_t s = rule_local_seconds(z.start, y) - z.std_off;
  const int64_t e = rule_local_seconds(z.end, y) - z.dst_off;
  const bool dst = s < e ? (utc >= s && utc < e) : !(utc >= e && utc < s);
  return dst ? z.dst_off : z.std_off;
}

bool load_file(const std::string& path, Zone& z) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  std::vector<unsigned char> buf;
  unsigned char tmp[65536];
  size_t n;
  while ((n = fread(tmp, 1, sizeof tmp, f)) > 0) {
    buf.insert(buf.end(), tmp, tmp + n);
    if (buf.size() > (8u << 20)) break;
  }
  fclose(f);
  if (buf.size() < 44 || memcmp(buf.data(), "TZif", 4) != 0) return false;
  const unsigned char* p = buf.data();
  const unsigned char* end = p + buf.size();
  const int version = p[4];
  auto counts = [&](const unsigned char* h, uint32_t c[6]) { for (int k = 0; k < 6; ++k) c[k] = (uint32_t)be32(h + 20 + 4 * k); };
  uint32_t c[6];  // isutcnt, isstdcnt, leapcnt, timecnt, typecnt, charcnt
  counts(p, c);
  size_t tsz = 4;  // width of a transition time in the block we read
  const unsigned char* d = p + 44;
  if (version >= '2') {
    const size_t v1 = (size_t)c[3] * 4 + c[3] + (size_t)c[4] * 6 + c[5] + (size_t)c[2] * 8 + c[1] + c[0];
    if (d + v1 + 44 > end || memcmp(d + v1, "TZif", 4) != 0) return false;
    p = d + v1;
    counts(p, c);
    d = p + 44;
    tsz = 8;
  }
  const size_t need = (size_t)c[3] * tsz + c[3] + (size_t)c[4] * 6 + c[5] + (size_t)c[2] * (tsz + 4) + c[1] + c[0];
  if (d + need > end || c[4] == 0) return false;
  const unsigned char* times = d;
  const unsigned char* idx = times + (size_t)c[3] * tsz;
  const unsigned char* types = idx + c[3];
  std::vector<int32_t> utoff(c[4]);
  std::vector<unsigned char> isdst(c[4]);
  for (uint32_t k = 0; k < c[4]; ++k) { utoff[k] = be32(types + 6 * k); isdst[k] = types[6 * k + 4]; }
  z.trans.resize(c[3]);
  z.off.resize(c[3]);
  for (uint32_t k = 0; k < c[3]; ++k) {
    z.trans[k] = tsz == 8 ? be64(times + 8 * k) : (int64_t)be32(times + 4 * k);
    if (idx[k] >= c[4]) return false;
    z.off[k] = utoff[idx[k]];
  }
  // before the first transition: the first standard-time type (as Go's lookupFirstZone), else type 0
  z.off_first = utoff[0];
  for (uint32_t k = 0; k < c[4]; ++k) if (!isdst[k]) { z.off_first = utoff[k]; break; }
  if (c[3] == 0) z.off_first = utoff[0];
  z.has_footer = false;
  if (tsz == 8) {
    const unsigned char* ft = d + need;
    if (ft < end && *ft == '\n') {
      const unsigned char* fe = (const unsigned char*)memchr(ft + 1, '\n', (size_t)(end - ft - 1));
      if (fe && fe > ft + 1) {
        std::string footer((const char*)ft + 1, (size_t)(fe - ft - 1));
        z.has_footer = parse_footer(footer, z);
      }
    }
  }
  return true;
}

int32_t zone_offset(const Zone& z, int64_t utc) {
  if (z.trans.empty() || utc < z.trans.front()) return z.trans.empty() && z.has_footer ? footer_offset(z, utc) : z.off_first;
  if (utc >= z.trans.back() && z.has_footer) return footer_offset(z, utc);
  size_t lo = 0, hi = z.trans.size();  // last transition <= utc
  while (hi - lo > 1) {
    const size_t mid = (lo + hi) / 2;
    if (z.trans[mid] <= utc) lo = mid; else hi = mid;
  }
  return z.off[lo];
}

std::mutex g_mu;
std::vector<std::unique_ptr<Zone>> g_zones;  // id - 1

bool valid_name(const std::string& n) {  // Go: containsDotDot, leading '/' or '\\' are "invalid location name"
  if (n.empty() || n[0] == '/' || n[0] == '\\' || n.size() > 255) return false;
  return n.find("..") == std::string::npos;
}

}  // namespace

int lookup(const char* name, size_t len, int32_t* id_out) {
  const std::string n(name ? name : "", len);
  if (n.empty() || n == "UTC" || n == "Local") { *id_out = 0; return 0; }  // time.Local == UTC in the shipped image
  if (!valid_name(n)) return -1;
  std::lock_guard<std::mutex> lk(g_mu);
  for (size_t k = 0; k < g_zones.size(); ++k)
    if (g_zones[k]->name == n) { *id_out = (int32_t)k + 1; return 0; }
  auto z = std::make_unique<Zone>();
  z->name = n;
  std::vector<std::string> dirs;
  if (const char* e = getenv("ZONEINFO")) dirs.emplace_back(e);
  dirs.insert(dirs.end(), {"/usr/share/zoneinfo", "/usr/share/lib/zoneinfo", "/usr/lib/locale/TZ", "/etc/zoneinfo"});
  bool ok = false;
  for (const std::string& d : dirs)
    if (load_file(d + "/" + n, *z)) { ok = true; break; }
  if (!ok) return -1;
  if ((int)g_zones.size() >= kMaxZones) return -2;
  g_zones.push_back(std::move(z));
  *id_out = (int32_t)g_zones.size();
  return 0;
}

bool offset_at(int32_t id, int64_t utc, int32_t* utoff) {
  if (id == 0) { *utoff = 0; return true; }
  std::lock_guard<std::mutex> lk(g_mu);
  if (id < 0 || (size_t)id > g_zones.size()) return false;
  *utoff = zone_offset(*g_zones[(size_t)id - 1], utc);
  return true;
}

int count() {
  std::lock_guard<std::mutex> lk(g_mu);
  return (int)g_zones.size();
}

bool tick_words(int64_t utc, amsweep::TickWords* table) {
  std::lock_guard<std::mutex> lk(g_mu);
  bool aligned = true;
  table[0] = amsweep::tick_words_from_unix(utc);
  for (size_t k = 0; k < g_zones.size(); ++k) {
    const int32_t off = zone_offset(*g_zones[k], utc);
    if (off % 60) aligned = false;
    table[k + 1] = amsweep::tick_words_from_unix(utc + off);
  }
  return aligned;
}

}  // namespace amsweep_tz
