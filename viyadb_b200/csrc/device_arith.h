// device_arith.h — the pure arithmetic of the device path: UTC calendar truncation and time rollup, the rank function of
// the time bucket dictionary, the hash mixers, the SmallerInt order image, the generator's random stream. No memory
// access, no intrinsics — so the very same source also compiles with plain g++ and is tested on the CPU
// (tests/device_arith_harness.cc, tests/test_device_arith.py) against the oracle. Under nvcc every function keeps the
// attributes it had inside kernels.cuh.
#ifndef VGPU_DEVICE_ARITH_H_
#define VGPU_DEVICE_ARITH_H_

#include "scan_params.h"

#if defined(__CUDACC__)
#define VGPU_HD __host__ __device__
#define VGPU_INLINE __forceinline__
#define VGPU_NOINLINE __noinline__
#else
#define VGPU_HD
#define VGPU_INLINE inline
#define VGPU_NOINLINE inline
#endif

namespace vgpu {

// ---------------------------------------------------------------------------------------------
// UTC calendar arithmetic (proleptic Gregorian, no leap seconds) == glibc gmtime_r / timegm for
// non-negative time_t, which is all util::Time32/Time64 ever see (unsigned inputs).
// ---------------------------------------------------------------------------------------------
// T = uint32_t for util::Time32 (seconds fit 32 bits: every division is by a constant, i.e. a multiply-high)
// and uint64_t for the seconds part of util::Time64.
template <typename T>
VGPU_HD VGPU_INLINE T trunc_days_to(T days, bool to_year) {
  // civil_from_days / days_from_civil (H. Hinnant's public-domain algorithms), days since 1970-01-01
  T z = days + 719468;
  T era = z / 146097;
  T doe = z - era * 146097;
  T yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
  T doy = doe - (365 * yoe + yoe / 4 - yoe / 100);  // March-based day of year
  T mp = (5 * doy + 2) / 153;
  T d = doy - (153 * mp + 2) / 5 + 1;
  if (!to_year) return days - (d - 1);
  // first of January of the civil year: March-based months 10,11 (Jan, Feb) belong to year yoe+1
  T y = yoe + era * 400 + (mp >= 10 ? 1 : 0);
  // days_from_civil(y, 1, 1)
  T yy = y - 1;
  T era2 = yy / 400;
  T yoe2 = yy - era2 * 400;
  T doy2 = (153 * 10 + 2) / 5;  // January 1st, March-based
  T doe2 = yoe2 * 365 + yoe2 / 4 - yoe2 / 100 + doy2;
  return era2 * 146097 + doe2 - 719468;
}

template <typename T>
VGPU_HD VGPU_INLINE T trunc_seconds(T t, uint32_t unit) {
  switch (unit) {
    case 0: return trunc_days_to(t / 86400, true) * 86400;   // YEAR
    case 1: return trunc_days_to(t / 86400, false) * 86400;  // MONTH
    case 3: return t - t % 86400;                            // DAY
    case 4: return t - t % 3600;                             // HOUR
    case 5: return t - t % 60;                               // MINUTE
    default: return t;                                       // SECOND / NONE
  }
}

VGPU_HD VGPU_NOINLINE uint64_t rollup_value(uint64_t v, const KeySpec &k) {
  uint32_t unit = 7;  // VGPU_TU_NONE
  for (uint32_t r = 0; r < k.nrules; ++r) {
    if (v < k.rule_boundary[r]) {  // first matching rule wins (rollup.cc:77-95)
      unit = k.rule_unit[r];
      break;
    }
  }
  // the query granularity truncates the same std::tm again (scan.cc:212-216): nested units, so the
  // result is the coarser of the two
  unit = unit < (uint32_t)k.query_unit ? unit : (uint32_t)k.query_unit;
  if (k.micro) {
    uint64_t secs = v / 1000000ull;
    uint64_t micros = v - secs * 1000000ull;
    if (unit != 7) micros = 0;  // Time64::trunc zeroes micros_ for every unit (time.h:129-132)
    return trunc_seconds<uint64_t>(secs, unit) * 1000000ull + micros;
  }
  return (uint64_t)trunc_seconds<uint32_t>((uint32_t)v, unit);
}

// rank of a rolled-up time value among all attainable ones (TimeDict, scan_params.h): no calendar arithmetic
VGPU_HD VGPU_NOINLINE uint64_t tdict_rank(const TimeDict &T, uint64_t v) {
  const uint64_t x = T.micro ? v / 1000000ull : v;
  // branch-free binary search for the last piece whose start is <= x (<= 48 pieces: 6 probes of the constant bank;
  // the linear scan it replaces was 27 % of the C4 kernel's instructions, profiles/r2_final_scan_ncu_c4.txt)
  uint32_t p = 0;
  if (T.narrow) {
    const uint32_t x32 = (uint32_t)x;
#pragma unroll
    for (uint32_t s = 32; s > 0; s >>= 1) {
      const uint32_t q = p + s;
      if (q < T.npieces && x32 >= T.start32[q]) p = q;
    }
  } else {
#pragma unroll
    for (uint32_t s = 32; s > 0; s >>= 1) {
      const uint32_t q = p + s;
      if (q < T.npieces && x >= T.start[q]) p = q;
    }
  }
  const uint32_t d = (uint32_t)(x - T.origin[p]);   // a piece spans less than 2^32 seconds
  uint32_t q;
  switch (T.step[p]) {   // divisions by compile-time constants
    case 0: q = 0; break;
    case 60: q = d / 60u; break;
    case 3600: q = d / 3600u; break;
    case 86400: q = d / 86400u; break;
    default: q = d; break;   // 1: second granularity
  }
  return (uint64_t)(T.base[p] + q);
}

// floating-point keys: -0.0 groups with +0.0 (KeyEqual uses ==, std::hash<float> maps both to 0; store.cc:46-85)
VGPU_HD VGPU_INLINE uint64_t fzero_fix(uint64_t val, uint32_t width) {
  return (width == 4 ? (uint32_t)(val << 1) == 0u : (val << 1) == 0ull) ? 0ull : val;
}

VGPU_HD VGPU_INLINE uint64_t mix64(uint64_t x) {
  x ^= x >> 33;
  x *= 0xff51afd7ed558ccdULL;
  x ^= x >> 33;
  x *= 0xc4ceb9fe1a85ec53ULL;
  x ^= x >> 33;
  return x;
}

// owner rank of a pair (scan kernel, several GPUs): every copy of a pair must meet on one rank
VGPU_HD VGPU_INLINE uint32_t pair_owner(uint64_t hi, uint64_t id, uint32_t nranks) {
  return (uint32_t)((mix64(hi * 0x9E3779B97F4A7C15ull + id) >> 33) % nranks);
}

VGPU_HD VGPU_INLINE uint32_t pair_bucket(uint64_t key, uint32_t nbuckets) {
  return (uint32_t)(mix64(key) >> 24) & (nbuckets - 1);
}

// Order-preserving image of an integer under util::StringNumCmp::SmallerInt (length, then lexicographic,
// src/util/string.h:28-49): non-negative numbers order numerically; "-" sorts before every digit, so inside one string
// length the negatives come first, by increasing magnitude. rank = number of representable values that sort before x.
VGPU_HD VGPU_INLINE uint64_t smaller_int_rank(long long x) {
  const unsigned long long p10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull,
                                      1000000000ull, 10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull,
                                      100000000000000ull, 1000000000000000ull, 10000000000000000ull, 100000000000000000ull,
                                      1000000000000000000ull, 10000000000000000000ull};
  const unsigned long long kPos = 1ull << 63;                 // representable non-negative values
  auto nonneg_upto = [&](int digits) { return digits <= 0 ? 0ull : (digits >= 19 ? kPos : p10[digits]); };      // with <= digits digits
  auto neg_upto = [&](int digits) { return digits <= 0 ? 0ull : (digits >= 19 ? kPos : p10[digits] - 1ull); };  // magnitude <= digits digits
  if (x >= 0) {
    const unsigned long long u = (unsigned long long)x;
    int d = 1;
    while (d < 19 && u >= p10[d]) ++d;                        // decimal digits = string length
    // before x: everything shorter (non-negatives with < d digits, negatives with magnitude < d - 1 digits ... of length < d),
    // the negatives of length d (magnitude of d - 1 digits), the non-negatives of d digits below x
    return nonneg_upto(d - 1) + neg_upto(d - 1) + (u - (d == 1 ? 0ull : p10[d - 1]));
  }
  const unsigned long long m = 0ull - (unsigned long long)x;  // magnitude (2^63 for INT64_MIN)
  int e = 1;
  while (e < 19 && m >= p10[e]) ++e;                          // digits of the magnitude; string length e + 1
  return nonneg_upto(e) + neg_upto(e - 1) + (m - p10[e - 1]);
}

// Order-preserving map of an element to uint64 (so one atomicMin/Max pair serves every type).
VGPU_HD VGPU_INLINE uint64_t to_ordered(uint64_t raw, uint32_t type) {
  switch (type) {
    case 4: case 5: case 6: case 7:  // signed ints (already sign-extended)
      return raw ^ 0x8000000000000000ull;
    case 8: {                        // f32
      uint32_t b = (uint32_t)raw;
      if (b == 0x80000000u) b = 0;   // -0.0 == +0.0
      b = (b >> 31) ? ~b : (b | 0x80000000u);
      return b;
    }
    case 9: {                        // f64
      uint64_t b = raw;
      if (b == 0x8000000000000000ull) b = 0;
      return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
    }
    default:
      return raw;
  }
}

VGPU_HD VGPU_INLINE uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// bit patterns as floating-point values (the kernel's intrinsics on the device, memcpy on the host)
#if defined(__CUDA_ARCH__)
VGPU_HD VGPU_INLINE float vgpu_bits_f32(uint32_t x) { return __uint_as_float(x); }
VGPU_HD VGPU_INLINE double vgpu_bits_f64(uint64_t x) { return __longlong_as_double((long long)x); }
#else
VGPU_HD VGPU_INLINE float vgpu_bits_f32(uint32_t x) { float f; __builtin_memcpy(&f, &x, 4); return f; }
VGPU_HD VGPU_INLINE double vgpu_bits_f64(uint64_t x) { double f; __builtin_memcpy(&f, &x, 8); return f; }
#endif

// ---------------------------------------------------------------------------------------------
// predicate leaves (filter.cc:206-261): the generic per-row compare and the vectorised leaf classes
// ---------------------------------------------------------------------------------------------
VGPU_HD VGPU_NOINLINE bool gen_compare(uint32_t gcls, uint32_t gop, uint64_t v, uint64_t a) {
  switch (gcls) {
    case G_I64: {
      long long x = (long long)v, y = (long long)a;
      switch (gop) { case 0: return x == y; case 1: return x != y; case 2: return x < y;
                     case 3: return x <= y; case 4: return x > y; default: return x >= y; }
    }
    case G_F32: {
      float x = vgpu_bits_f32((uint32_t)v), y = vgpu_bits_f32((uint32_t)a);
      switch (gop) { case 0: return x == y; case 1: return x != y; case 2: return x < y;
                     case 3: return x <= y; case 4: return x > y; default: return x >= y; }
    }
    case G_F64: {
      double x = vgpu_bits_f64(v), y = vgpu_bits_f64(a);
      switch (gop) { case 0: return x == y; case 1: return x != y; case 2: return x < y;
                     case 3: return x <= y; case 4: return x > y; default: return x >= y; }
    }
    default: {  // G_U64, G_CARD
      switch (gop) { case 0: return v == a; case 1: return v != a; case 2: return v < a;
                     case 3: return v <= a; case 4: return v > a; default: return v >= a; }
    }
  }
}

// mask of one vectorisable leaf over the 16 rows in v
VGPU_HD VGPU_INLINE uint32_t leaf_mask16(const PInstr &in, const uint32_t (&v)[kRowsPerThread]) {
  uint32_t m = 0;
  const uint32_t cls = in.cls;
  const uint32_t a = (uint32_t)in.arg;
  if (cls == C_EQ32) {
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) if (v[i] == a) m |= 1u << i;
  } else if (cls == C_LT32) {
    const uint32_t bias = in.bias;
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) if ((v[i] ^ bias) < a) m |= 1u << i;
  } else if (cls == C_RNG32) {
    const uint32_t bias = in.bias, len = in.arg2;
    if (bias == 0) {
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) if ((v[i] - a) < len) m |= 1u << i;
    } else {
#pragma unroll
      for (int i = 0; i < kRowsPerThread; ++i) if (((v[i] ^ bias) - a) < len) m |= 1u << i;
    }
  } else {  // C_LUT64: membership in a set of codes < 64 — one shift per row whatever the list length
    const uint64_t lut = in.arg;
#pragma unroll
    for (int i = 0; i < kRowsPerThread; ++i) {
      uint64_t t;
#if defined(__CUDA_ARCH__)
      asm("shr.b64 %0, %1, %2;" : "=l"(t) : "l"(lut), "r"(v[i]));  // shift amounts >= 64 give 0
#else
      t = v[i] >= 64u ? 0ull : lut >> v[i];
#endif
      if (t & 1ull) m |= 1u << i;
    }
  }
  return m;
}

// HAVING on the device (post_agg.cc:76-83): generic compare of widened group values (group_passes, kernels.cuh)
VGPU_HD VGPU_INLINE bool post_compare(uint32_t gcls, uint32_t gop, uint64_t v, uint64_t a) {
  switch (gcls) {
    case G_I64: {
      const long long x = (long long)v, y = (long long)a;
      switch (gop) { case 0: return x == y; case 1: return x != y; case 2: return x < y; case 3: return x <= y; case 4: return x > y; default: return x >= y; }
    }
    case G_F32: {
      const float x = vgpu_bits_f32((uint32_t)v), y = vgpu_bits_f32((uint32_t)a);
      switch (gop) { case 0: return x == y; case 1: return x != y; case 2: return x < y; case 3: return x <= y; case 4: return x > y; default: return x >= y; }
    }
    case G_F64: {
      const double x = vgpu_bits_f64(v), y = vgpu_bits_f64(a);
      switch (gop) { case 0: return x == y; case 1: return x != y; case 2: return x < y; case 3: return x <= y; case 4: return x > y; default: return x >= y; }
    }
    default:
      switch (gop) { case 0: return v == a; case 1: return v != a; case 2: return v < a; case 3: return v <= a; case 4: return v > a; default: return v >= a; }
  }
}

}  // namespace vgpu

#endif  // VGPU_DEVICE_ARITH_H_
