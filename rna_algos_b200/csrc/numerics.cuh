// numerics.cuh — device versions of the reference's approximate log-sum-exp / exp
// (reference src/utils.rs:579-655).  Bit-exact by construction: every product and sum is a separate
// IEEE f32 operation (__fmul_rn/__fadd_rn never contract to FMA), evaluated in the reference's order,
// with the reference's f32 coefficient literals.
#pragma once
#include "portable.h"

namespace rna {

#define RNA_NEG_INF (__int_as_float(0xff800000))
// RNA_FASTNUM = 1: the FAST_F32 numeric mode's build of the batch kernel (csrc/fold_fastnum.cu) — same phases, roles
// and data layout, but logsumexp / exp in exact log-space arithmetic (MUFU ex2 / lg2) instead of the reference's
// piecewise polynomials, and the long chains accumulate in linear space (fold_phases.cuh ChainSum).
#ifndef RNA_FASTNUM
#define RNA_FASTNUM 0
#endif
#define RNA_LSE_THRESHOLD 11.862479f

// Coefficient table of ln_exp_1p: 8 segments x {a, b, c, d}, poly = ((a*x + b)*x + c)*x + d.
// Staged into shared memory once per CTA (per-lane segment index => LDS.128, conflict-free: each
// 16-byte row lives in its own 4 banks, equal rows broadcast).
RNA_CONST_TABLE float4 kLnExp1pCoef[8] = {
    {-0.0065591595f, 0.12764427f, 0.49965546f, 0.6931542f},     // x < 0.66153675
    {-0.015515756f, 0.14467756f, 0.48829398f, 0.6958093f},      // x < 1.6320158
    {-0.012890925f, 0.13010283f, 0.51503986f, 0.6795586f},      // x < 2.4912589
    {-0.0072142647f, 0.087754086f, 0.6208708f, 0.5909676f},     // x < 3.37925
    {-0.0031455354f, 0.046722945f, 0.7592532f, 0.43487945f},    // x < 4.426169
    {-0.0010110698f, 0.018594341f, 0.88317305f, 0.25236955f},   // x < 5.789071
    {-0.000196278f, 0.0046084408f, 0.9634432f, 0.09831489f},    // x < 7.8162727
    {-0.0000113994f, 0.0003734731f, 0.9959107f, 0.0149855051f}  // else
};

#ifdef __CUDACC__
__device__ __forceinline__ void load_lse_lut(float4* lut_smem) {
  if (threadIdx.x < 8) lut_smem[threadIdx.x] = kLnExp1pCoef[threadIdx.x];
}
#endif

// Segment index of the reference's comparison tree (src/utils.rs:604-626), branch-free.
RNA_DEV int ln_exp_1p_segment(float x) {
  const bool p1 = x < 3.37925f;
  const float t2 = p1 ? 1.6320158f : 5.789071f;
  const bool p2 = x < t2;
  const float t3 = p1 ? (p2 ? 0.66153675f : 2.4912589f) : (p2 ? 4.426169f : 7.8162727f);
  const bool p3 = x < t3;
  return (p1 ? 0 : 4) + (p2 ? 0 : 2) + (p3 ? 0 : 1);
}

RNA_DEV float ln_exp_1p(float x, const float4* __restrict__ lut) {
  const float4 c = lut[ln_exp_1p_segment(x)];
  float r = __fadd_rn(__fmul_rn(c.x, x), c.y);
  r = __fadd_rn(__fmul_rn(r, x), c.z);
  r = __fadd_rn(__fmul_rn(r, x), c.w);
  return r;
}

RNA_DEV bool is_finite(float x) { return fabsf(x) < __int_as_float(0x7f800000); }

// logsumexp(&mut sum, x): src/utils.rs:580-596.  Returns the new sum.
// The reference ignores a non-finite x and replaces a non-finite sum by x.  Here x is first normalised off the
// critical path (NaN -> -inf; +inf never occurs), after which ONE select on the chain covers every case:
//   both finite        z = max - min is finite            -> y + (z >= T ? z : poly(z))
//   exactly one -inf   z = +inf                           -> max(sum, x) = the finite one   (bitwise: fmaxf picks it)
//   both -inf          z = NaN, (z < inf) is false        -> max = -inf
RNA_DEV float lse(float sum, float x, const float4* __restrict__ lut) {
  x = (x > RNA_NEG_INF) ? x : RNA_NEG_INF;
#if RNA_FASTNUM
  {   // exact log-space arithmetic on the MUFU pipe: max + log(1 + exp(min - max)); both -inf stays -inf
    const float mx = fmaxf(sum, x);
    const float r = __fadd_rn(mx, __logf(__fadd_rn(1.f, __expf(__fsub_rn(fminf(sum, x), mx)))));
    return (mx > RNA_NEG_INF) ? r : RNA_NEG_INF;
  }
#endif
  const float y = fminf(sum, x);
  const float mx = fmaxf(sum, x);
  const float z = __fsub_rn(mx, y);
  // z is NaN/inf when an operand is -inf; the LUT index stays in range (all compares false => 7).
  const float r = ln_exp_1p(z, lut);
  const float v = __fadd_rn(y, (z >= RNA_LSE_THRESHOLD) ? z : r);
  return (z < __int_as_float(0x7f800000)) ? v : mx;
}

// Latency-optimised logsumexp, bit-identical to lse(): no shared-memory LUT — the seven breakpoint compares issue side
// by side right after z, and each coefficient comes out of a three-level select tree (only the tree of `a` is on the
// dependency chain; b, c, d resolve under the Horner steps).  42 instructions instead of 26, but a dependent step
// takes 94 cycles instead of 114 on a warp that has its scheduler to itself (tools/lse_microbench.cu): used by the
// cooperative long-sequence kernel, whose chains are pure latency; the batch kernels, which are issue-bound, keep lse().
RNA_DEV float lse_sel8(bool q1, bool q2, bool q3, bool q4, bool q5, bool q6, bool q7, float c0, float c1, float c2,
                       float c3, float c4, float c5, float c6, float c7) {
  const float lo = q2 ? (q3 ? c3 : c2) : (q1 ? c1 : c0);
  const float hi = q6 ? (q7 ? c7 : c6) : (q5 ? c5 : c4);
  return q4 ? hi : lo;
}
// NORM = false: the caller guarantees that x is finite or -inf (never NaN), which saves the normalising select
template <bool NORM = true>
RNA_DEV float lse_lat(float sum, float x) {
  if (NORM) x = (x > RNA_NEG_INF) ? x : RNA_NEG_INF;
  const float y = fminf(sum, x);
  const float mx = fmaxf(sum, x);
  const float z = __fsub_rn(mx, y);
  // !(z < b): a NaN z lands in the last segment like in the reference's comparison tree; its result is discarded below
  const bool q1 = !(z < 0.66153675f), q2 = !(z < 1.6320158f), q3 = !(z < 2.4912589f), q4 = !(z < 3.37925f),
             q5 = !(z < 4.426169f), q6 = !(z < 5.789071f), q7 = !(z < 7.8162727f);
  const float a = lse_sel8(q1, q2, q3, q4, q5, q6, q7, -0.0065591595f, -0.015515756f, -0.012890925f, -0.0072142647f,
                           -0.0031455354f, -0.0010110698f, -0.000196278f, -0.0000113994f);
  const float b = lse_sel8(q1, q2, q3, q4, q5, q6, q7, 0.12764427f, 0.14467756f, 0.13010283f, 0.087754086f,
                           0.046722945f, 0.018594341f, 0.0046084408f, 0.0003734731f);
  const float c = lse_sel8(q1, q2, q3, q4, q5, q6, q7, 0.49965546f, 0.48829398f, 0.51503986f, 0.6208708f, 0.7592532f,
                           0.88317305f, 0.9634432f, 0.9959107f);
  const float d = lse_sel8(q1, q2, q3, q4, q5, q6, q7, 0.6931542f, 0.6958093f, 0.6795586f, 0.5909676f, 0.43487945f,
                           0.25236955f, 0.09831489f, 0.0149855051f);
  float r = __fadd_rn(__fmul_rn(a, z), b);
  r = __fadd_rn(__fmul_rn(r, z), c);
  r = __fadd_rn(__fmul_rn(r, z), d);
  const float v = __fadd_rn(y, (z >= RNA_LSE_THRESHOLD) ? z : r);
  return (z < __int_as_float(0x7f800000)) ? v : mx;
}

// logsumexp into an EMPTY sum (sum == -inf): the reference takes x if it is finite (src/utils.rs:581-586)
RNA_DEV float lse_init(float x) { return (x > RNA_NEG_INF) ? x : RNA_NEG_INF; }

// Running sum of a long chain (the two-loop streams, the dense split-point chains, the rightmost-pair chains).
// Reference-exact build: the log-space value itself, folded with lse() term by term in the reference's order.
// FAST build: a LINEAR-space accumulator relative to a reference exponent — value = ref + log(s) — so a step is one
// ex2 off the dependency chain plus one dependent FADD; the reference moves only when a term exceeds it by e^60
// (that also handles the first finite term: ref = -inf), so s stays far inside the f32 range.
#if RNA_FASTNUM
struct ChainSum { float s, ref; };
RNA_DEV ChainSum chain_begin(float sumlog) {
  ChainSum c;
  c.ref = sumlog;
  c.s = (sumlog > RNA_NEG_INF) ? 1.f : 0.f;
  return c;
}
RNA_DEV void chain_add(ChainSum& c, float x, const float4* __restrict__) {
  // branch-free.  t = +inf: the first finite term (ref = -inf); t = NaN: -inf - -inf or a NaN operand (skipped)
  const float t = __fsub_rn(x, c.ref);
  const bool lead = t > 60.f;
  const float e = __expf(lead ? -t : t);
  const float sn = lead ? fmaf(c.s, e, 1.f) : __fadd_rn(c.s, e);
  c.s = (t == t) ? sn : c.s;
  c.ref = lead ? x : c.ref;
}
RNA_DEV float chain_end(const ChainSum& c) { return (c.s > 0.f) ? __fadd_rn(c.ref, __logf(c.s)) : RNA_NEG_INF; }
#else
typedef float ChainSum;
RNA_DEV float chain_begin(float sumlog) { return sumlog; }
RNA_DEV void chain_add(float& c, float x, const float4* __restrict__ lut) { c = lse(c, x, lut); }
RNA_DEV float chain_end(float c) { return c; }
#endif

// lse with a per-lane enable predicate (disabled lanes keep `sum`).
RNA_DEV float lse_if(bool on, float sum, float x, const float4* __restrict__ lut) {
  return lse(sum, on ? x : RNA_NEG_INF, lut);
}

// expf: src/utils.rs:631-655.  For x >= 0 the reference calls libm expf; here exp is evaluated in f64
// and rounded once to f32, which equals the correctly rounded f32 result (and glibc's <1-ULP expf)
// except for ~1e-8 of inputs (DESIGN.md "numerics").
RNA_DEV float approx_expf(float x) {
#if RNA_FASTNUM
  return expf(x);
#endif
  if (x < -2.4915035f) {
    if (x < -5.8622823f) {
      if (x < -9.91152f) return 0.f;
      return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(0.0000803850f, x), 0.002162743f), x), 0.019470856f), x), 0.058808003f);
    } else if (x < -3.839663f) {
      return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(0.0013889414f, x), 0.024467647f), x), 0.14712906f), x), 0.30427578f);
    } else {
      return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(0.0072335607f, x), 0.09060027f), x), 0.39831114f), x), 0.62459594f);
    }
  } else if (x < -0.6725053f) {
    if (x < -1.4805375f) {
      return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(0.023241036f, x), 0.2085646f), x), 0.6906368f), x), 0.86823225f);
    } else {
      return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(0.057378277f, x), 0.35802585f), x), 0.9121133f), x), 0.9793092f);
    }
  } else if (x < 0.f) {
    return __fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(0.119917594f, x), 0.48156682f), x), 0.9975992f), x), 0.9999505f);
  }
  return (float)exp((double)x);
}

RNA_DEV bool canonical_pair(int x, int y) {
  // AU CG GC GU UA UG with A=0 C=1 G=2 U=3: bitmask over x*4+y
  return (0x5a48u >> (x * 4 + y)) & 1u;   // bits 3(AU) 6(CG) 9(GC) 11(GU) 12(UA) 14(UG)
}
RNA_DEV bool augu_pair(int x, int y) {
  return (0x5808u >> (x * 4 + y)) & 1u;   // bits 3(AU) 11(GU) 12(UA) 14(UG)
}

}  // namespace rna
