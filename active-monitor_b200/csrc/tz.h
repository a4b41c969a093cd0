// tz.h — named time zones for cron schedules ("CRON_TZ=Europe/Paris 0 9 * * *").
//
// robfig/cron v3.0.1 (parser.go, call site hcc.go:253) resolves the prefix with
// time.LoadLocation and evaluates the schedule in that zone.  Round 1 rejected such specs
// (AM_E_UNSUPPORTED, kind HOST_FALLBACK) and nothing evaluated them.  Now: a process-wide
// registry maps zone names to small ids (carried in bits 24..31 of a record's flags); per tick
// the host computes, per registered zone, the broken-down LOCAL time of the tick as one-hot
// words (the same TickWords the kernel ANDs against the cron masks) and uploads the 256-entry
// table; the kernel picks table[tz_id] for a zone-bound record — still five ANDs, no loop.
//
// Zone rules come from the system's TZif files (RFC 8536; $ZONEINFO, /usr/share/zoneinfo, ... —
// the directories Go's time.LoadLocation searches): the 64-bit transition table plus the POSIX TZ
// footer for instants past the last transition (slim files carry little else).  No libc, no
// global TZ state: the oracle uses libc / zoneinfo instead, so the two share nothing.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include <vector>

#include "civil.h"
#include "tz_eval.h"

namespace amsweep_tz {

constexpr int kMaxZones = 255;  // ids 1..255; 0 = UTC / time.Local of the shipped image

// Register (or find) a zone by name.  0 ok (*id_out set; "", "UTC", "Local" -> 0); -1 the zone
// cannot be loaded (robfig: "provided bad location"); -2 registry full.
int lookup(const char* name, size_t len, int32_t* id_out);
// UTC offset (seconds east) of zone `id` at a UTC instant.  false: unknown id.
bool offset_at(int32_t id, int64_t utc, int32_t* utoff);
int count();
// Flattened copy of the registry for the device (entry 0 = UTC) and its version (bumped on every
// registration); descs == nullptr: the version only.
uint64_t snapshot(std::vector<ZoneDesc>* descs, std::vector<int64_t>* trans, std::vector<int32_t>* off);
// true iff every registered zone's offset at `utc` is a whole number of minutes: the kernel may then
// skip the cron masks off the minute (the local second-of-minute equals the UTC one)
bool all_minute_aligned(int64_t utc);
// The tick's local broken-down time per zone (entry 0 = UTC).  Returns true iff every registered
// zone's offset at `utc` is a whole number of minutes (the kernel may then skip the cron masks off
// the minute: the local second-of-minute equals the UTC one).
bool tick_words(int64_t utc, amsweep::TickWords* table /* [kMaxZones + 1] */);

}  // namespace amsweep_tz
