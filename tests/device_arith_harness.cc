// device_arith_harness.cc — the device path's pure arithmetic (viyadb_b200/csrc/device_arith.h) and the host side of the
// time bucket dictionary (csrc/time_dict.h), compiled with plain g++ and exported over a C ABI for the CPU tests
// (tests/test_device_arith.py). TEST INFRASTRUCTURE: the very source the CUDA kernels compile, run on the host — it is
// not a CPU path of the product (libvgpu.so has none) and nothing under viyadb_b200/ links it.
#include "../viyadb_b200/csrc/device_arith.h"
#include "../viyadb_b200/csrc/time_dict.h"

#include <cstring>

using namespace vgpu;

extern "C" {

// trunc_seconds<uint32_t> (util::Time32) or <uint64_t> (seconds part of util::Time64)
void h_trunc_seconds(const uint64_t *t, uint64_t n, uint32_t unit, int wide, uint64_t *out) {
  for (uint64_t i = 0; i < n; ++i)
    out[i] = wide ? trunc_seconds<uint64_t>(t[i], unit) : (uint64_t)trunc_seconds<uint32_t>((uint32_t)t[i], unit);
}

// util::Truncator as the reference runs it: glibc gmtime_r -> clear fields -> timegm (host_trunc, time_dict.h)
void h_glibc_trunc(const uint64_t *t, uint64_t n, uint32_t unit, uint64_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = unit >= VGPU_TU_SECOND ? t[i] : host_trunc(t[i], unit);
}

static void fill_key(KeySpec &k, int micro, uint32_t nrules, const uint8_t *units, const uint64_t *boundaries, uint32_t query_unit) {
  std::memset(&k, 0, sizeof k);
  k.rollup = 1;
  k.micro = micro ? 1 : 0;
  k.nrules = (uint8_t)nrules;
  for (uint32_t r = 0; r < nrules; ++r) {
    k.rule_unit[r] = units[r];
    k.rule_boundary[r] = boundaries[r];
  }
  k.query_unit = (uint8_t)query_unit;
}

// rollup_value: what the scan kernel computes per passing row when no dictionary applies
void h_rollup(const uint64_t *v, uint64_t n, int micro, uint32_t nrules, const uint8_t *units, const uint64_t *boundaries,
              uint32_t query_unit, uint64_t *out) {
  KeySpec k;
  fill_key(k, micro, nrules, units, boundaries, query_unit);
  for (uint64_t i = 0; i < n; ++i) out[i] = rollup_value(v[i], k);
}

// build_time_dict (planner, host) + tdict_rank (kernel) over the same rules. Returns 0 when the planner declines the
// dictionary, else the number of bucket values written to values[] (capacity max_values; -1: too many).
long long h_time_dict(const uint64_t *v, uint64_t n, int micro, uint32_t nrules, const uint8_t *units, const uint64_t *boundaries,
                      uint32_t query_unit, uint64_t raw_lo, uint64_t raw_hi, uint64_t *ranks, uint64_t *values,
                      uint64_t max_values, uint32_t *npieces, uint32_t *narrow) {
  vgpu_key key;
  std::memset(&key, 0, sizeof key);
  key.nrules = nrules;
  for (uint32_t r = 0; r < nrules; ++r) {
    key.rule_granularity[r] = units[r];
    key.rule_boundary[r] = boundaries[r];
  }
  key.query_granularity = query_unit;
  TimeDict T;
  std::vector<uint64_t> vals;
  if (!build_time_dict(key, micro != 0, raw_lo, raw_hi, T, vals)) return 0;
  if (vals.size() > max_values) return -1;
  for (size_t i = 0; i < vals.size(); ++i) values[i] = vals[i];
  for (uint64_t i = 0; i < n; ++i) ranks[i] = tdict_rank(T, v[i]);
  *npieces = T.npieces;
  *narrow = T.narrow;
  return (long long)vals.size();
}

void h_smaller_int_rank(const long long *x, uint64_t n, uint64_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = smaller_int_rank(x[i]);
}

void h_to_ordered(const uint64_t *raw, uint64_t n, uint32_t type, uint64_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = to_ordered(raw[i], type);
}

void h_fzero_fix(const uint64_t *raw, uint64_t n, uint32_t width, uint64_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = fzero_fix(raw[i], width);
}

void h_splitmix64(const uint64_t *x, uint64_t n, uint64_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = splitmix64(x[i]);
}

void h_mix64(const uint64_t *x, uint64_t n, uint64_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = mix64(x[i]);
}

// owner rank of a count-distinct pair (multi-GPU exchange) and its hash bucket (dedupe): every copy of a pair must land
// in one place, and the places must be balanced
void h_pair_owner(const uint64_t *hi, const uint64_t *id, uint64_t n, uint32_t nranks, uint32_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = pair_owner(hi[i], id[i], nranks);
}
void h_pair_bucket(const uint64_t *key, uint64_t n, uint32_t nbuckets, uint32_t *out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = pair_bucket(key[i], nbuckets);
}

}  // extern "C"
