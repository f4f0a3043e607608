// time_dict.h — host side of the time bucket dictionary (TimeDict, scan_params.h): built once per query from the rule
// boundaries with gmtime_r / timegm, exactly the calls util::Truncator makes (src/util/time.h:52-89). Plain C++ (no
// CUDA): included by vgpu.cu for the planner and by tests/device_arith_harness.cc for the CPU tests.
#ifndef VGPU_TIME_DICT_H_
#define VGPU_TIME_DICT_H_

#include "../../include/vgpu.h"
#include "scan_params.h"

#include <algorithm>
#include <stdint.h>
#include <time.h>
#include <vector>

namespace vgpu {

// ---------------------------------------------------------------------------------------------
// bucket dictionary of a rolled-up time key (TimeDict, scan_params.h)
// ---------------------------------------------------------------------------------------------
// util::Truncator (src/util/time.h:52-89) on the host: gmtime_r / timegm exactly like the reference
inline uint64_t host_trunc(uint64_t secs, uint32_t unit) {
  time_t tt = (time_t)secs;
  struct tm tm;
  gmtime_r(&tt, &tm);
  if (unit <= VGPU_TU_YEAR) tm.tm_mon = 0;
  if (unit <= VGPU_TU_MONTH) tm.tm_mday = 1;
  if (unit <= VGPU_TU_DAY) tm.tm_hour = 0;
  if (unit <= VGPU_TU_HOUR) tm.tm_min = 0;
  if (unit <= VGPU_TU_MINUTE) tm.tm_sec = 0;
  return (uint64_t)timegm(&tm);
}
inline uint64_t host_next_period(uint64_t start, uint32_t unit) {  // start of the month / year after the one beginning at `start`
  time_t tt = (time_t)start;
  struct tm tm;
  gmtime_r(&tt, &tm);
  if (unit == VGPU_TU_YEAR) tm.tm_year += 1; else tm.tm_mon += 1;
  return (uint64_t)timegm(&tm);
}

// false: no dictionary (the kernel rolls every row up with calendar arithmetic, the key keeps its raw domain)
inline bool build_time_dict(const vgpu_key &key, bool micro, uint64_t raw_lo, uint64_t raw_hi, TimeDict &T,
                     std::vector<uint64_t> &values) {
  const uint64_t scale = micro ? 1000000ull : 1ull;
  const uint64_t xs_lo = raw_lo / scale, xs_hi = raw_hi / scale;
  if (xs_hi - xs_lo >= 0xffffffffull) return false;
  struct Piece { uint64_t sel, origin; uint32_t step; uint64_t nb; };
  std::vector<Piece> pieces;
  // regions of the raw time line: rule r applies to [largest earlier boundary, boundary r) — "first rule with
  // value < boundary wins" (rollup.cc:77-95); past the last boundary only the query granularity truncates
  uint64_t prev = 0;  // in seconds
  for (uint32_t r = 0; r <= key.nrules; ++r) {
    uint64_t end;  // exclusive, seconds
    uint32_t unit = VGPU_TU_NONE;
    if (r < key.nrules) {
      const uint64_t b = key.rule_boundary[r];
      if (micro && b % scale != 0 && b > raw_lo && b <= raw_hi) return false;  // boundary inside a second
      end = micro ? (b + scale - 1) / scale : b;
      unit = key.rule_granularity[r];
      if (end <= prev) continue;  // shadowed by an earlier rule
    } else {
      end = ~0ull;
    }
    unit = std::min<uint32_t>(unit, key.query_granularity);  // nested units: the coarser of the two wins
    const uint64_t a = std::max(prev, xs_lo), b = std::min<uint64_t>(end, xs_hi + 1);
    prev = end == ~0ull ? prev : end;
    if (a >= b) continue;
    if (unit == VGPU_TU_WEEK) return false;
    if (unit >= VGPU_TU_SECOND) {
      if (micro && unit == VGPU_TU_NONE) return false;  // Time64 keeps the microseconds when nothing truncates
      pieces.push_back({a, a, 1, b - a});
    } else if (unit >= VGPU_TU_DAY) {
      const uint32_t st = unit == VGPU_TU_DAY ? 86400u : unit == VGPU_TU_HOUR ? 3600u : 60u;
      const uint64_t origin = a - a % st;
      pieces.push_back({a, origin, st, (b - 1 - origin) / st + 1});
    } else {
      for (uint64_t p = host_trunc(a, unit); p < b; p = host_next_period(p, unit)) {
        pieces.push_back({std::max(p, a), p, 0, 1});
        if (pieces.size() > (size_t)kMaxTimeSegs) return false;
      }
    }
    if (pieces.size() > (size_t)kMaxTimeSegs) return false;
  }
  if (pieces.empty()) return false;
  T = TimeDict{};
  values.clear();
  uint64_t last_value = 0, last_rank = 0;
  bool have = false;
  for (size_t j = 0; j < pieces.size(); ++j) {
    const Piece &pc = pieces[j];
    uint64_t base;
    if (!have) base = 0;
    else if (pc.origin == last_value) base = last_rank;  // the same truncated value on both sides of a boundary
    else if (pc.origin > last_value) base = last_rank + 1;
    else return false;                                    // not monotone: a finer rule before a coarser one
    last_value = pc.origin + (pc.nb - 1) * pc.step;
    last_rank = base + pc.nb - 1;
    have = true;
    if (last_rank >= (1ull << 22)) return false;
    for (uint64_t q = 0; q < pc.nb; ++q)
      if (base + q >= values.size()) values.push_back((pc.origin + q * pc.step) * scale);
    T.start[j] = pc.sel;
    T.origin[j] = pc.origin;
    T.base[j] = (uint32_t)base;
    T.step[j] = pc.step;
  }
  T.npieces = (uint32_t)pieces.size();
  T.micro = micro ? 1u : 0u;
  T.start[0] = 0;  // whatever the statistics say, piece 0 catches every smaller value
  T.narrow = xs_hi <= 0xffffffffull ? 1u : 0u;
  for (uint32_t j = 0; j < T.npieces; ++j) {
    if (T.start[j] > 0xffffffffull) T.narrow = 0;
    T.start32[j] = (uint32_t)T.start[j];
  }
  return true;
}

}  // namespace vgpu

#endif  // VGPU_TIME_DICT_H_
