// scorers.cuh — Turner 2004 / CONTRAfold v2.02 loop scorers (reference src/utils.rs:166-556) over the
// packed device tables.  Portable (see portable.h): compiled by nvcc for the kernels and by g++ for the
// host emulator in tests/emu.
#pragma once
#include "dev_tables.h"
#include "numerics.cuh"

namespace rna {

RNA_DEV int idx4(int a, int b, int c, int d) { return ((a * 4 + b) * 4 + c) * 4 + d; }
RNA_DEV int idx3(int a, int b, int c) { return (a * 4 + b) * 4 + c; }

// ---------------------------------------------------------------------------------------------------
// Turner 2004 scorers (src/utils.rs:166-411).  `s` = sequence bytes (shared memory), T = staged tables.
// ---------------------------------------------------------------------------------------------------
struct TurnerView {
  const DevTurner* g;        // global (uniform reads)
  const TurnerSmall* sm;     // shared copy (per-lane gathers)
};

template <class TV>
RNA_DEV float t_pen(const TV& T, int x, int y) {
  return augu_pair(x, y) ? T.g->augu_pen : 0.f;
}

// get_hairpin_score, src/utils.rs:166-196 (special-loop scan :198-205 via packed keys)
template <class TV>
RNA_DEV float t_hairpin(const TV& T, const uint8_t* s, int i, int j) {
  const int span = j - i + 1;
  if (span < 32 && ((T.g->special_len_mask >> span) & 1u)) {
    unsigned key = 0;
    for (int p = 0; p < span; p++) key |= (unsigned)s[i + p] << (2 * p);
    const int n = T.g->num_special;
    for (int x = 0; x < n; x++) {
      if ((int)T.g->special_len[x] == span && T.g->special_key[x] == key) {
        // first match decides (src/utils.rs:198-205); a listed score of -inf falls through to the generic
        // hairpin (`if special_hairpin_score > NEG_INFINITY`, src/utils.rs:169)
        if (T.g->special_score[x] > RNA_NEG_INF) return T.g->special_score[x];
        break;
      }
    }
  }
  const int len = j - i - 1;
  const int si = s[i], sj = s[j];
  float hs;
  if (len == T.g->min_hairpin_len) {
    hs = __ldg(&T.g->hairpin_init_ext[len]);
  } else {
    hs = __fadd_rn(__ldg(&T.g->hairpin_init_ext[len]), T.sm->tm_hairpin[idx4(si, sj, s[i + 1], s[j - 1])]);
  }
  return __fadd_rn(hs, t_pen(T, si, sj));
}

// get_2loop_score, src/utils.rs:207-366.  (i,j) closes, (k,l) is enclosed; a = k-i-1, b = j-l-1 (warp-uniform).
RNA_DEV float t_twoloop(const TurnerView& T, const uint8_t* s, int i, int j, int k, int l,
                                           int a, int b) {
  const int si = s[i], sj = s[j], sk = s[k], sl = s[l];
  if (a == 0 && b == 0) return T.sm->stack[idx4(si, sj, sk, sl)];
  if (a == 0 || b == 0) {
    const int len = a + b;
    const float bi = T.g->bulge_init[len];
    if (len == 1) return __fadd_rn(bi, T.sm->stack[idx4(si, sj, sk, sl)]);
    return __fadd_rn(__fadd_rn(bi, t_pen(T, si, sj)), t_pen(T, sk, sl));
  }
  if (a <= 2 && b <= 2) {
    const int i1 = s[i + 1], j1 = s[j - 1];
    if (a == 1 && b == 1) return __ldg(&T.g->int11[idx4(si, sj, i1, j1) * 16 + sk * 4 + sl]);
    if (a == 1 && b == 2) return __ldg(&T.g->int12[(idx4(si, sj, i1, j1) * 4 + s[j - 2]) * 16 + sk * 4 + sl]);
    if (a == 2 && b == 1)
      return __ldg(&T.g->int12[(idx4(sl, sk, j1, s[i + 2]) * 4 + i1) * 16 + sj * 4 + si]);
    return __ldg(&T.g->int22[(idx4(si, sj, i1, j1) * 16 + s[i + 2] * 4 + s[j - 2]) * 16 + sk * 4 + sl]);
  }
  const float* tm = (a == 1 || b == 1) ? T.sm->tm_1xmany
                    : ((a == 2 && b == 3) || (a == 3 && b == 2)) ? T.sm->tm_2x3 : T.sm->tm_interior;
  const float mm = __fadd_rn(tm[idx4(si, sj, s[i + 1], s[j - 1])], tm[idx4(sl, sk, s[l + 1], s[k - 1])]);
  float v = __fadd_rn(T.g->interior_init_ninio[a * 31 + b], mm);
  v = __fadd_rn(v, t_pen(T, si, sj));
  return __fadd_rn(v, t_pen(T, sk, sl));
}

// get_multibranch_close_score, src/utils.rs:368-382
template <class TV>
RNA_DEV float t_mbclose(const TV& T, const uint8_t* s, int i, int j) {
  const int si = s[i], sj = s[j];
  const float tm = T.sm->tm_multi[idx4(sj, si, s[j - 1], s[i + 1])];
  return __fadd_rn(__fadd_rn(T.g->init_mb_base, tm), t_pen(T, si, sj));
}

// get_accessible_score (uses_sentinel_bases = false), src/utils.rs:384-411
template <class TV>
RNA_DEV float t_acc(const TV& T, const uint8_t* s, int L, int i, int j) {
  const int si = s[i], sj = s[j];
  float sc;
  if (i > 0 && j < L - 1) sc = T.sm->tm_multi[idx4(si, sj, s[i - 1], s[j + 1])];
  else if (i > 0) sc = T.sm->d5[idx3(si, sj, s[i - 1])];
  else if (j < L - 1) sc = T.sm->d3[idx3(si, sj, s[j + 1])];
  else sc = 0.f;
  return __fadd_rn(sc, t_pen(T, si, sj));
}

// ---------------------------------------------------------------------------------------------------
// CONTRAfold v2.02 scorers (src/utils.rs:413-556)
// ---------------------------------------------------------------------------------------------------
struct ContraView {
  const DevContra* g;
  const ContraSmall* sm;
};

// get_junction_score_single, src/utils.rs:545-556
RNA_DEV float c_jsingle(const ContraView& T, const uint8_t* s, int p0, int p1) {
  const int x = s[p0], y = s[p1];
  return __fadd_rn(T.sm->hc[x * 4 + y], T.sm->tm[idx4(x, y, s[p0 + 1], s[p1 - 1])]);
}

// get_junction_score (uses_sentinel_bases = false), src/utils.rs:522-543
RNA_DEV float c_junction(const ContraView& T, const uint8_t* s, int L, int p0, int p1) {
  const int x = s[p0], y = s[p1];
  float v = __fadd_rn(T.sm->hc[x * 4 + y], (p0 < L - 1) ? T.sm->dl[idx3(x, y, s[min(p0 + 1, L - 1)])] : 0.f);
  return __fadd_rn(v, (p1 > 0) ? T.sm->dr[idx3(x, y, s[max(p1 - 1, 0)])] : 0.f);
}

// get_hairpin_score_contra, src/utils.rs:413-421
RNA_DEV float c_hairpin(const ContraView& T, const uint8_t* s, int i, int j) {
  const int len = j - i - 1;
  return __fadd_rn(T.g->hairpin_cum[min(len, T.g->max_loop_len)], c_jsingle(T, s, i, j));
}

// get_2loop_score_contra, src/utils.rs:423-520
RNA_DEV float c_twoloop(const ContraView& T, const uint8_t* s, int i, int j, int k, int l,
                                           int a, int b) {
  const int sk = s[k], sl = s[l];
  float sc;
  if (a == 0 && b == 0) {
    sc = T.sm->stack[idx4(s[i], s[j], sk, sl)];
  } else if (a == 0 || b == 0) {
    const int len = a + b;
    const float s0 = (len == 1) ? T.sm->bulge0x1[(a == 1) ? s[i + 1] : s[j - 1]] : 0.f;
    sc = __fadd_rn(s0, T.g->bulge_cum[len - 1]);
    sc = __fadd_rn(sc, c_jsingle(T, s, i, j));
    sc = __fadd_rn(sc, c_jsingle(T, s, l, k));
  } else {
    const int len = a + b;
    float v;
    if (a == b) {
      const float s11 = (len == 2) ? T.sm->int1x1[s[i + 1] * 4 + s[j - 1]] : 0.f;
      v = __fadd_rn(s11, T.g->sym_cum[a - 1]);
    } else {
      v = T.g->asym_cum[(a > b ? a - b : b - a) - 1];
    }
    const float ex = (a <= T.g->max_explicit && b <= T.g->max_explicit) ? T.g->explicit_[(a - 1) * 4 + (b - 1)] : 0.f;
    v = __fadd_rn(v, ex);
    v = __fadd_rn(v, T.g->interior_cum[len - 2]);
    v = __fadd_rn(v, c_jsingle(T, s, i, j));
    sc = __fadd_rn(v, c_jsingle(T, s, l, k));
  }
  return __fadd_rn(sc, T.sm->bp[sk * 4 + sl]);
}

template <bool CONTRA> struct ModelTraits;
template <> struct ModelTraits<false> { typedef DevTurner Dev; typedef TurnerSmall Small; typedef TurnerView View; };
template <> struct ModelTraits<true> { typedef DevContra Dev; typedef ContraSmall Small; typedef ContraView View; };

template <bool CONTRA>
RNA_DEV float m_twoloop(const typename ModelTraits<CONTRA>::View& T, const uint8_t* s, int i,
                                           int j, int k, int l, int a, int b) {
  if constexpr (CONTRA) return c_twoloop(T, s, i, j, k, l, a, b);
  else return t_twoloop(T, s, i, j, k, l, a, b);
}
template <bool CONTRA>
RNA_DEV float m_mbclose(const typename ModelTraits<CONTRA>::View& T, const uint8_t* s, int L,
                                           int i, int j) {
  if constexpr (CONTRA) return __fadd_rn(T.g->mb_base_plus_bp, c_junction(T, s, L, i, j));
  else return t_mbclose(T, s, i, j);
}
template <bool CONTRA>
RNA_DEV float m_acc(const typename ModelTraits<CONTRA>::View& T, const uint8_t* s, int L, int i,
                                       int j) {
  if constexpr (CONTRA) return __fadd_rn(c_junction(T, s, L, j, i), T.sm->bp[s[i] * 4 + s[j]]);
  else return t_acc(T, s, L, i, j);
}

}  // namespace rna
