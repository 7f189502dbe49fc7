// Host-side exact arithmetic of the SEA worst-case mIoU search.
//
// evalSEA.worst_case_miou (tools/worse_only.py:267-334) is a sequential greedy: for every image,
// in a shuffled order, and every attack it scores "what if this image took that attack" as
// statistics.mean over the classes of (run_int + d_int) / (run_union + d_union + 1e-8) and keeps
// the reassignment when the score drops below the current mIoU.  statistics.mean is the
// CORRECTLY ROUNDED mean of the exact rational sum, and the running sums pass through float32
// every time the reference rebuilds a tensor from its lists (:311-316,323-326), so a bit-identical
// result needs exactly that arithmetic.  The reference spends minutes in .item() loops here at
// N = 2000 images; this is the same sequence of operations in C++ (one call per round; the
// shuffles stay with Python's `random` so the seeded order is the reference's).
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace robseg {
namespace {

// Exact accumulator for finite doubles: unsigned fixed point with the least significant bit at
// 2^-1126 (below half the smallest subnormal's last bit) in 36 64-bit limbs.
struct BigSum {
  static constexpr int kLimbs = 36;
  static constexpr int kBias = 1126;
  uint64_t w[kLimbs];
  BigSum() { std::memset(w, 0, sizeof(w)); }

  void add_magnitude(double x) {  // x > 0, finite
    int e;
    const double m = std::frexp(x, &e);                       // x = m * 2^e, m in [0.5, 1)
    const uint64_t mant = (uint64_t)std::ldexp(m, 53);        // 53-bit integer mantissa
    const int pos = e - 53 + kBias;                           // bit position of mant's LSB (>= 0)
    const int limb = pos >> 6, sh = pos & 63;
    uint64_t lo = mant << sh, hi = sh ? (mant >> (64 - sh)) : 0;
    uint64_t c = 0;
    uint64_t t = w[limb] + lo;
    c = t < lo;
    w[limb] = t;
    t = w[limb + 1] + hi;
    uint64_t c2 = t < hi;
    t += c;
    c2 |= t < c;
    w[limb + 1] = t;
    for (int i = limb + 2; c2 && i < kLimbs; ++i) c2 = (++w[i] == 0);
  }
  static int cmp(const BigSum& a, const BigSum& b) {
    for (int i = kLimbs - 1; i >= 0; --i)
      if (a.w[i] != b.w[i]) return a.w[i] < b.w[i] ? -1 : 1;
    return 0;
  }
  void sub(const BigSum& b) {  // *this >= b
    uint64_t borrow = 0;
    for (int i = 0; i < kLimbs; ++i) {
      const uint64_t bi = b.w[i] + borrow;
      const uint64_t nb = (bi < borrow) || (w[i] < bi);
      w[i] -= bi;
      borrow = nb;
    }
  }
  // round-half-even(value / n) as a double, value = this * 2^-kBias
  double div_to_double(uint64_t n) const {
    uint64_t q[kLimbs];
    unsigned __int128 rem = 0;
    for (int i = kLimbs - 1; i >= 0; --i) {
      const unsigned __int128 cur = (rem << 64) | w[i];
      q[i] = (uint64_t)(cur / n);
      rem = cur % n;
    }
    int top = kLimbs - 1;
    while (top >= 0 && q[top] == 0) --top;
    if (top < 0) return 0.0;
    const int hb = top * 64 + 63 - __builtin_clzll(q[top]);  // highest set bit
    // 64 bits starting at hb (zero-extended below bit 0), everything lower folded into `sticky`
    uint64_t bits = 0;
    bool sticky = rem != 0;
    const int lo_bit = hb - 63;
    if (lo_bit <= 0) {
      bits = q[0] << (-lo_bit);  // hb < 64: the whole quotient fits
    } else {
      const int l = lo_bit >> 6, s = lo_bit & 63;
      bits = q[l] >> s;
      if (s) bits |= q[l + 1] << (64 - s);
      if (s && (q[l] & ((1ull << s) - 1))) sticky = true;
      for (int i = 0; i < l; ++i) sticky |= q[i] != 0;
    }
    uint64_t mant = bits >> 11;           // 53 bits
    const uint64_t low = bits & 0x7ff;    // 11 guard bits
    if (low > 0x400 || (low == 0x400 && (sticky || (mant & 1)))) ++mant;  // 2^53 is still exact below
    return std::ldexp((double)mant, hb - 52 - kBias);
  }
};

// statistics.mean(values): correctly rounded exact_sum / n.  false if a value is not finite.
bool exact_mean(const double* v, int64_t n, double* out) {
  BigSum pos, neg;
  for (int64_t i = 0; i < n; ++i) {
    const double x = v[i];
    if (!std::isfinite(x)) return false;
    if (x > 0) pos.add_magnitude(x);
    else if (x < 0) neg.add_magnitude(-x);
  }
  if (BigSum::cmp(pos, neg) >= 0) {
    pos.sub(neg);
    *out = pos.div_to_double((uint64_t)n);
  } else {
    neg.sub(pos);
    *out = -neg.div_to_double((uint64_t)n);
  }
  return true;
}

inline double f32(double x) { return (double)(float)x; }  // torch.tensor(list) -> float32 (:311-312)

}  // namespace
}  // namespace robseg

using namespace robseg;

extern "C" int robseg_exact_mean_host(const double* values_host, int64_t n, double* mean_host) {
  ROBSEG_REQUIRE(values_host && mean_host && n > 0, "mean requires at least one data point");
  ROBSEG_REQUIRE(exact_mean(values_host, n, mean_host), "non-finite value");
  return 0;
}

extern "C" int robseg_sea_greedy_round_host(const double* cons_ints_host, const double* cons_unions_host,
                                            int A, int N, int C, const int32_t* order_host,
                                            int32_t* sel_host, double* run_int_host,
                                            double* run_union_host, double* final_miou_host) {
  ROBSEG_REQUIRE(cons_ints_host && cons_unions_host && order_host && sel_host && run_int_host &&
                     run_union_host && final_miou_host,
                 "NULL pointer");
  ROBSEG_REQUIRE(A > 0 && N > 0 && C > 0, "bad shape");
  const double* ci = cons_ints_host;
  const double* cu = cons_unions_host;
  std::vector<double> ri(C), ru(C), ni(C), nu(C), ratio(C);
  double final_miou = *final_miou_host;
  for (int t = 0; t < N; ++t) {
    const int idx = order_host[t];
    ROBSEG_REQUIRE(idx >= 0 && idx < N, "order[%d]=%d out of range", t, idx);
    for (int a = 0; a < A; ++a) {
      const int s = sel_host[idx];
      ROBSEG_REQUIRE(s >= 0 && s < A, "sel[%d]=%d out of range", idx, s);
      const double* ia = ci + ((size_t)a * N + idx) * C;
      const double* is = ci + ((size_t)s * N + idx) * C;
      const double* ua = cu + ((size_t)a * N + idx) * C;
      const double* us = cu + ((size_t)s * N + idx) * C;
      int k = 0;
      for (int c = 0; c < C; ++c) {
        ri[c] = f32(run_int_host[c]), ru[c] = f32(run_union_host[c]);
        ni[c] = ri[c] + (ia[c] - is[c]);
        nu[c] = ru[c] + (ua[c] - us[c]);
        if (ru[c] != 0) ratio[k++] = ni[c] / (nu[c] + 1e-8);  // candidate score (:79-93)
      }
      ROBSEG_REQUIRE(k > 0, "every running union is zero");
      double est;
      ROBSEG_REQUIRE(exact_mean(ratio.data(), k, &est), "non-finite candidate score");
      if (est < final_miou) {
        sel_host[idx] = a;
        std::memcpy(run_int_host, ni.data(), sizeof(double) * C);
        std::memcpy(run_union_host, nu.data(), sizeof(double) * C);
      }
    }
    int k = 0;
    for (int c = 0; c < C; ++c) {
      const double i32 = f32(run_int_host[c]), u32 = f32(run_union_host[c]);
      if (u32 != 0) ratio[k++] = i32 / u32;
    }
    ROBSEG_REQUIRE(k > 0, "every running union is zero");
    ROBSEG_REQUIRE(exact_mean(ratio.data(), k, &final_miou), "non-finite mIoU");
  }
  *final_miou_host = final_miou;
  return 0;
}
