// fold_phases.cuh — the McCaskill inside / outside recurrences of one sequence, cut into "phases": the work
// one role of the CTA does for one anti-diagonal between two barriers.  Portable (portable.h): nvcc compiles
// the phases into fold_kernel2.cuh's kernel; g++ compiles the same text into the host emulator in tests/emu,
// which runs the phases barrier by barrier to check the parallel decomposition against the oracle.
//
// Reference recurrences: src/mccaskill_algo.rs:282-378 (Turner inside), :380-516 (CONTRAfold inside),
// :518-610 / :612-723 (outside + BPP).  Every cell's fold is evaluated in the reference's order with the
// reference's polynomial logsumexp, so all values are bit-identical to the reference algorithm.
//
// What differs from a thread-per-cell wavefront (fold_kernel.cuh):
//   * closability is static (canonical pair + span rule), so it is a per-sequence BIT MATRIX built once;
//     interior-loop partners (k,l) of a cell are found by scanning 31-bit windows of row k with CLZ/FFS:
//     work is proportional to the number of real terms instead of 496 probes per cell;
//   * the two-loop chains run one lane per CLOSABLE cell (compacted per-diagonal lists), each lane walking
//     its own partner list, so no lane idles on a non-pairable cell;
//   * the chains of one diagonal are split over warp ROLES that run concurrently:
//       inside step t :  X = sums_close(t)   |  Y = rightmost-pair sums (t) over closable (i,k), k<j
//                        Z = finish R/Rm(t-1), then sums_external / sums_multibranch / sums_1ormore (t-1)
//       outside step t:  X = P(i,j) for closable cells (exterior + enclosing two-loops + multiloops)
//                        Y = probs_multibranch / probs_multibranch2 for all cells
//     one barrier per step; dependencies are argued at each phase.
#pragma once
#include "scorers.cuh"

namespace rna {

struct ModelParams {
  int MAX2;          // MAX_2LOOP_LEN (Turner) / MAX_LOOP_LEN (CONTRAfold)
  int MINSPAN;       // MIN_SPAN_HAIRPIN_CLOSE
  int allows_short;  // CONTRAfold only
};

// One sequence's working set.  Matrices are DIAGONAL-MAJOR: index(i,j) = doff(j-i) + i.
// PIdx = element type of the closable-cell lists (uint8_t when L <= 256, else uint16_t).
template <class PIdx>
struct SeqViewT {
  int L;
  int W2;                 // words per bit-matrix row: ceil(L/32) data words + one zero pad word at each end
  const uint8_t* s;       // bases; s[-4..-1] and s[L..L+3] are readable zeros
  uint32_t* mask;         // [L][W2]: bit l of row k (data words start at index 1) <=> (k,l) statically closable
  PIdx* plist;            // diagonal-major: plist[doff(d) + r] = i of the r-th closable cell of diagonal d
  uint16_t* pcnt;         // [L] closable cells per diagonal
  uint8_t* RR;            // [L] base codes s[p]*4 + s[p+1]   (s[L] = 0)
  uint8_t* LL;            // [L] base codes s[p]*4 + s[p-1]   (s[-1] = 0)
  // static two-loop term streams (global memory, see "term streams" below); tin == null => scored on the fly
  int din0, dout0;        // first diagonal of the inside pass / last diagonal of the outside pass
  // A STEP of a pass handles two diagonals: inside step s = (din0+2s, din0+2s+1), outside step s = (L-1-2s, L-2-2s).
  // The closable cells of a step, diagonal A first, are cut into lane GROUPS of 32 (the lanes of one warp).
  uint32_t* gcumI;        // [NS+1] groups before inside step s
  uint32_t* gcumO;        // [NS+1] groups before outside step s
  uint32_t* ccumI;        // [NS+1] closable cells before inside step s
  uint32_t* ccumO;        // [NS+1] closable cells before outside step s
  uint16_t* gstepI;       // [NG] step of an inside group
  uint16_t* gstepO;       // [NG] step of an outside group
  uint32_t* gbin;         // [NG+1] first element of each inside group's block in tin
  uint32_t* gbout;        // [NG+1] ... of each outside group's block in tout
  uint2* tin;             // closing-pair-major: terms of the inside chains, lane-interleaved per group
  uint2* tout;            // enclosed-pair-major: terms of the outside chains
  uint16_t* ccnt;         // scratch [2][L(L+1)/2]: terms per closable cell (indexed like plist)
  uint32_t tcap;          // capacity of tin / tout in elements
  float* C;               // sums_close
  float* R;               // sums_rightmost_basepairs_external   -> outside: probs_multibranch
  float* X;               // CONTRAfold: sums_rightmost_basepairs_multibranch -> outside: probs_multibranch2
  float* E;               // sums_external
  float* Pm;              // outside: log P(i,j) -> BPP  (may alias E: sums_external is dead after the inside pass)
  float* M1;              // sums_1ormore_basepairs
  float* Mroll;           // sums_multibranch, 3 rolling diagonals of L
  float* E0;              // sums_external[0][x]
  float* EL;              // sums_external[x][L-1]
  // cooperative kernel, outside pass only (null elsewhere): ROW-MAJOR triangular copies, index(i,j) = doff(i) + j-i
  float* M1rm;            // sums_1ormore_basepairs, row-major
  float* MB;              // multibranch closing score of every closable (i,j) (0 elsewhere), diagonal-major
  // optional export of the whole sums_multibranch matrix (only three diagonals are live otherwise): the
  // RNA_SUMS_MULTIBRANCH plane of rna_fold_sums_batch, row-major upper triangle incl. the diagonal; null = off
  float* Mfull;
};

// index of (i,j), i <= j, in a plane of rna_fold_sums_batch (include/rna_algos_b200.h: rna_sums_index)
RNA_DEV size_t sums_index(int L, int i, int j) { return (size_t)i * L - ((size_t)i * (i - 1)) / 2 + (size_t)(j - i); }


RNA_DEV int doff(int d, int L) { return d * L - ((d * (d - 1)) >> 1); }

// The chains of the cooperative kernel are pure latency (one warp per scheduler): they fold with the select-tree
// logsumexp (numerics.cuh lse_lat, bit-identical to lse).  -DRNA_COOP_LSE_LUT switches back for A/B timing.
#ifdef RNA_COOP_LSE_LUT
#define RNA_COOP_LSE(sum, x, lut) lse(sum, x, lut)
#else
#define RNA_COOP_LSE(sum, x, lut) lse_lat(sum, x)
#endif
// the same for operands that are sums of finite or -inf values (no subtraction, no +inf): they cannot be NaN
#ifdef RNA_COOP_LSE_LUT
#define RNA_COOP_LSE_NN(sum, x, lut) lse(sum, x, lut)
#else
#define RNA_COOP_LSE_NN(sum, x, lut) lse_lat<false>(sum, x)
#endif
// (see "The long chains of the cooperative kernel ..." below)
#ifdef __CUDACC__
#define RNA_CHAIN_FN __device__ __noinline__
#define RNA_CHAIN_PTR(name) __device__ decltype(&name) g_##name = name;
#define RNA_CHAIN_CALL(name) (*reinterpret_cast<decltype(&name) volatile*>(&g_##name))
#else
#define RNA_CHAIN_FN static inline
#define RNA_CHAIN_PTR(name)
#define RNA_CHAIN_CALL(name) name
#endif

// The cooperative long-sequence kernel's view: same fields, but sums_close / log P are far away (HBM/L2).
struct CoopView : SeqViewT<uint16_t> {};

// bits [pos, pos+31] of a bit-matrix row (`row` points at the leading pad word); pos in [-32, 32*(W2-2))
RNA_DEV uint32_t get32(const uint32_t* row, int pos) {
  const int q = (pos >> 5) + 1;
  return __funnelshift_r(row[q], row[q + 1], (unsigned)(pos & 31));
}

template <bool CONTRA>
RNA_DEV bool closable_static(const uint8_t* s, int i, int j, const ModelParams& P) {
  if (!canonical_pair(s[i], s[j])) return false;
  if (CONTRA && P.allows_short) return true;
  return j - i + 1 >= P.MINSPAN;
}

// ---- setup ---------------------------------------------------------------------------------------------
template <bool CONTRA, class SV>
RNA_DEV void setup_mask_word(const SV& v, const ModelParams& P, int x) {
  const int k = x / v.W2, wi = x - k * v.W2;
  uint32_t bits = 0;
  if (wi >= 1 && wi <= v.W2 - 2) {
    const int base = (wi - 1) * 32;
    for (int t = 0; t < 32; t++) {
      const int l = base + t;
      if (l < v.L && l > k && closable_static<CONTRA>(v.s, k, l, P)) bits |= 1u << t;
    }
  }
  v.mask[x] = bits;
}

// the two diagonals of a step and their closable-cell counts
struct StepCells { int dA, dB, cA, cB; };
template <bool INSIDE, class SV>
RNA_DEV StepCells step_cells(const SV& v, int st) {
  StepCells c;
  if (INSIDE) {
    c.dA = v.din0 + 2 * st; c.dB = c.dA + 1;
    c.cA = (c.dA < v.L) ? v.pcnt[c.dA] : 0;
    c.cB = (c.dB < v.L) ? v.pcnt[c.dB] : 0;
  } else {
    c.dA = v.L - 1 - 2 * st; c.dB = c.dA - 1;
    c.cA = (c.dA >= v.dout0) ? v.pcnt[c.dA] : 0;
    c.cB = (c.dB >= v.dout0) ? v.pcnt[c.dB] : 0;
  }
  return c;
}
template <class SV> RNA_DEV int num_steps_inside(const SV& v) { return (v.L - v.din0 + 1) / 2; }
template <class SV> RNA_DEV int num_steps_outside(const SV& v) { return (v.L - v.dout0 + 1) / 2; }
// x-th cell of a step -> (diagonal, rank on the diagonal)
RNA_DEV void step_cell(const StepCells& c, int x, int& dd, int& r) {
  const bool first = x < c.cA;
  dd = first ? c.dA : c.dB;
  r = first ? x : x - c.cA;
}
RNA_DEV uint32_t group_width(int tot, int c) { return (uint32_t)min(32, tot - 32 * c); }

// group bases of every step (serial over <= L entries: one thread)
template <class SV>
RNA_DEV void setup_groups(const SV& v) {
  uint32_t run = 0, cells = 0;
  const int nsi = max(num_steps_inside(v), 0), nso = max(num_steps_outside(v), 0);
  for (int st = 0; st < nsi; st++) {
    const StepCells c = step_cells<true>(v, st);
    v.gcumI[st] = run;
    v.ccumI[st] = cells;
    const uint32_t ng = (uint32_t)(c.cA + c.cB + 31) >> 5;
    for (uint32_t g = 0; g < ng; g++) v.gstepI[run + g] = (uint16_t)st;
    run += ng;
    cells += (uint32_t)(c.cA + c.cB);
  }
  v.gcumI[nsi] = run;
  v.ccumI[nsi] = cells;
  run = 0; cells = 0;
  for (int st = 0; st < nso; st++) {
    const StepCells c = step_cells<false>(v, st);
    v.gcumO[st] = run;
    v.ccumO[st] = cells;
    const uint32_t ng = (uint32_t)(c.cA + c.cB + 31) >> 5;
    for (uint32_t g = 0; g < ng; g++) v.gstepO[run + g] = (uint16_t)st;
    run += ng;
    cells += (uint32_t)(c.cA + c.cB);
  }
  v.gcumO[nso] = run;
  v.ccumO[nso] = cells;
}

template <class SV>
RNA_DEV void setup_codes(const SV& v, int p) {
  const uint8_t* s = v.s;   // s[-1] and s[L] are zero pads
  v.RR[p] = (uint8_t)(s[p] * 4 + s[p + 1]);
  v.LL[p] = (uint8_t)(s[p] * 4 + s[p - 1]);
}

template <class PIdx>
RNA_DEV void setup_list_diag(const SeqViewT<PIdx>& v, int d) {
  const int L = v.L, od = doff(d, L);
  int cnt = 0;
  for (int i = 0; i + d < L; i++) {
    const int j = i + d;
    if ((v.mask[i * v.W2 + 1 + (j >> 5)] >> (j & 31)) & 1u) v.plist[od + cnt++] = (PIdx)i;
  }
  v.pcnt[d] = (uint16_t)cnt;
}

// ---- v2 views -----------------------------------------------------------------------------------------------
struct ContraView2 {
  const DevContra* g;
  const ContraSmall2* sm;
};
struct TurnerView2 {
  const DevTurner* g;
  const TurnerSmall2* sm;
};
RNA_DEV float c2_js(const ContraView2& T, const uint8_t* s, int p0, int p1) {
  return T.sm->js2[(s[p0] * 4 + s[p0 + 1]) * 16 + s[p1] * 4 + s[p1 - 1]];
}
RNA_DEV float c2_junction(const ContraView2& T, const uint8_t* s, int L, int p0, int p1) {
  const int x = s[p0], y = s[p1];
  float v = __fadd_rn(T.sm->hc[x * 4 + y], (p0 < L - 1) ? T.sm->dl[idx3(x, y, s[min(p0 + 1, L - 1)])] : 0.f);
  return __fadd_rn(v, (p1 > 0) ? T.sm->dr[idx3(x, y, s[max(p1 - 1, 0)])] : 0.f);
}
RNA_DEV float c2_hairpin(const ContraView2& T, const uint8_t* s, int i, int j) {
  return __fadd_rn(T.g->hairpin_cum[min(j - i - 1, T.g->max_loop_len)], c2_js(T, s, i, j));
}

template <bool CONTRA> struct Model2;
template <> struct Model2<false> {
  typedef DevTurner Dev; typedef TurnerSmall2 Small; typedef TurnerView2 View;
};
template <> struct Model2<true> {
  typedef DevContra Dev; typedef ContraSmall2 Small; typedef ContraView2 View;
};
template <bool CONTRA>
RNA_DEV const typename Model2<CONTRA>::Small* dev_small(const typename Model2<CONTRA>::Dev* d) {
  return &d->small2;
}

template <bool CONTRA>
RNA_DEV float v2_mbclose(const typename Model2<CONTRA>::View& T, const uint8_t* s, int L, int i, int j) {
  if constexpr (CONTRA) return __fadd_rn(T.g->mb_base_plus_bp, c2_junction(T, s, L, i, j));
  else return t_mbclose(T, s, i, j);
}
template <bool CONTRA>
RNA_DEV float v2_acc(const typename Model2<CONTRA>::View& T, const uint8_t* s, int L, int i, int j) {
  if constexpr (CONTRA) return __fadd_rn(c2_junction(T, s, L, j, i), T.sm->bp[s[i] * 4 + s[j]]);
  else return t_acc(T, s, L, i, j);
}

// =========================================================================================================
// Two-loop chains (src/mccaskill_algo.rs:305-323, 574-593 / 410-434, 681-700): one lane folds, in the
// reference's order, over the partner pairs (k,l) of its cell (i,j).  The chain of logsumexp's is strictly
// sequential, so its latency IS the critical path of a diagonal; everything that does not depend on the
// running sum is taken off it by a 3-stage software pipeline over the lane's partner list:
//     stage 1 (term n+2)  advance the bit-window iterator; first-level loads: C[q] (and log P[q]), base code
//     stage 2 (term n+1)  second-level gathers: the score-table entries addressed by the base code
//     stage 3 (term n)    combine the loaded values into the operand y; sum = logsumexp(sum, y)
// An exhausted iterator keeps feeding c = -inf, which makes y = -inf and the fold a no-op (utils.rs:580-583).
//
// Base codes: RR[p] = s[p]*4+s[p+1], LL[p] = s[p]*4+s[p-1].  A partner's code is RR[5' base]*16 + LL[3' base]
// of the side that VARIES: INSIDE the enclosed pair seen from outside in, (l,k): RR[l]*16+LL[k];
// OUTSIDE the closing pair (k,l): RR[k]*16+LL[l].  code = [x:2][x1:2][y:2][y1:2].
// =========================================================================================================
struct Term1 { float c, pv; int code, a, b, q; };     // q = matrix index of the partner, -1 = none
struct Term2 { float c, pv, u, v; int cls, aux, q; };

// ---- CONTRAfold: get_2loop_score_contra, src/utils.rs:423-520 -------------------------------------------------
template <bool INSIDE>
struct ContraLoop {
  const ContraView2& T;
  float js_fixed, bp_fixed;   // INSIDE: js(i,j), unused.  OUTSIDE: js(j,i), basepair_scores[s[i]][s[j]]
  int pq, p1, q1;             // s[i]*4+s[j]; INSIDE: s[i+1], s[j-1]
  RNA_DEVM ContraLoop(const ContraView2& T_, const uint8_t* RR, const uint8_t* LL, int i, int j) : T(T_) {
    const int ri = RR[i], lj = LL[j];
    pq = (ri >> 2) * 4 + (lj >> 2);
    p1 = ri & 3;
    q1 = lj & 3;
    if (INSIDE) { js_fixed = T.sm->js2[ri * 16 + lj]; bp_fixed = 0.f; }
    else { js_fixed = T.sm->js2[RR[j] * 16 + LL[i]]; bp_fixed = T.sm->bp[pq]; }
  }
  RNA_DEVM Term2 stage2(const Term1& t) const {
    Term2 r;
    r.c = t.c; r.pv = t.pv; r.q = t.q;
    const int a = t.a, b = t.b, code = t.code;
    const int x = code >> 6, x1 = (code >> 4) & 3, y = (code >> 2) & 3, y1 = code & 3;
    const bool st = (a | b) == 0;
    int ui;
    if (INSIDE) {   // code = (l,k): x = s[l], y = s[k]
      ui = st ? pq * 16 + y * 4 + x
              : (a + b == 1) ? RNA_CU_B1 + (a == 1 ? p1 : q1)
              : (a == 1 && b == 1) ? RNA_CU_I11 + p1 * 4 + q1 : RNA_CU_PTAB + a * 31 + b;
      r.aux = y * 4 + x;   // s[k]*4+s[l]
    } else {        // code = (k,l) closing: x = s[k], x1 = s[k+1], y = s[l], y1 = s[l-1]
      ui = st ? (x * 4 + y) * 16 + pq
              : (a + b == 1) ? RNA_CU_B1 + (a == 1 ? x1 : y1)
              : (a == 1 && b == 1) ? RNA_CU_I11 + x1 * 4 + y1 : RNA_CU_PTAB + a * 31 + b;
      r.aux = 0;
    }
    r.u = T.sm->U[ui];
    r.v = T.sm->js2[code];
    r.cls = st ? 1 : 0;
    return r;
  }
  // rows a >= 2 hold no stack / bulge-of-1 / 1x1 term: one gather of the length table, same additions
  static constexpr int kFastFrom = 2;
  RNA_DEVM float fast(int a, int b, int code) const {
    const float u = T.sm->U[RNA_CU_PTAB + a * 31 + b], vv = T.sm->js2[code];
    if (INSIDE) {
      const float sc = __fadd_rn(__fadd_rn(u, js_fixed), vv);
      return __fadd_rn(sc, T.sm->bp[((code >> 2) & 3) * 4 + (code >> 6)]);
    }
    return __fadd_rn(__fadd_rn(__fadd_rn(u, vv), js_fixed), bp_fixed);
  }
  // the two-loop score: a pure function of the sequence (the reference memoises it: mccaskill_algo.rs:431,696)
  RNA_DEVM float score(const Term2& t) const {
    if (INSIDE) {
      const float sc = t.cls ? t.u : __fadd_rn(__fadd_rn(t.u, js_fixed), t.v);
      return __fadd_rn(sc, T.sm->bp[t.aux]);
    } else {
      const float sc = t.cls ? t.u : __fadd_rn(__fadd_rn(t.u, t.v), js_fixed);
      return __fadd_rn(sc, bp_fixed);
    }
  }
};

// ---- Turner: get_2loop_score, src/utils.rs:207-366 ------------------------------------------------------------
template <bool INSIDE>
struct TurnerLoop {
  const TurnerView2& T;
  const uint8_t* s;
  int i, j;
  float tm_f0, tm_f1, tm_f2;   // interior-mismatch term of the fixed pair for the three table classes
  float pen_fixed;     // AU/GU end penalty of the fixed pair
  int pq;              // s[i]*4+s[j]
  RNA_DEVM TurnerLoop(const TurnerView2& T_, const uint8_t* s_, const uint8_t* RR, const uint8_t* LL, int i_, int j_)
      : T(T_), s(s_), i(i_), j(j_) {
    const int si = s[i], sj = s[j];
    pq = si * 4 + sj;
    pen_fixed = t_pen(T, si, sj);
    const int code = INSIDE ? RR[i] * 16 + LL[j] : RR[j] * 16 + LL[i];
    tm_f0 = T.sm->tm2[0][code];
    tm_f1 = T.sm->tm2[1][code];
    tm_f2 = T.sm->tm2[2][code];
  }
  // cls: 0 stack, 1 bulge of 1, 2 longer bulge, 3 explicit 1x1/1x2/2x1/2x2 table, 4.. generic interior (4 + table class)
  RNA_DEVM Term2 stage2(const Term1& t) const {
    Term2 r;
    r.c = t.c; r.pv = t.pv; r.q = t.q; r.u = 0.f; r.v = 0.f;
    const int a = t.a, b = t.b, code = t.code;
    const int x = code >> 6, x1 = (code >> 4) & 3, y = (code >> 2) & 3, y1 = code & 3;
    // INSIDE: closing (i,j) fixed, enclosed (k,l): x = s[l], x1 = s[l+1], y = s[k], y1 = s[k-1]
    // OUTSIDE: enclosed (i,j) fixed, closing (k,l): x = s[k], x1 = s[k+1], y = s[l], y1 = s[l-1]
    const int si = INSIDE ? (pq >> 2) : x, sj = INSIDE ? (pq & 3) : y;   // closing pair
    const int sk = INSIDE ? y : (pq >> 2), sl = INSIDE ? x : (pq & 3);   // enclosed pair
    r.aux = augu_pair(INSIDE ? sk : si, INSIDE ? sl : sj) ? 1 : 0;       // AU/GU penalty of the varying pair
    const int st = (si * 4 + sj) * 16 + sk * 4 + sl;
    if ((a | b) == 0) {
      r.cls = 0;
      r.u = T.sm->stack[st];
    } else if (a == 0 || b == 0) {
      const int len = a + b;
      r.v = T.sm->bulge_init[len];
      if (len == 1) { r.cls = 1; r.u = T.sm->stack[st]; } else { r.cls = 2; }
    } else if (a <= 2 && b <= 2) {
      r.cls = 3;
      // closing pair (ci,cj) and its inner neighbours i1 = s[ci+1], j1 = s[cj-1], i2 = s[ci+2], j2 = s[cj-2]
      int i1, j1, i2, j2;
      if (INSIDE) { i1 = s[i + 1]; j1 = s[j - 1]; i2 = s[i + 2]; j2 = s[j - 2]; }
      else {
        const int k = i - 1 - a, l = j + 1 + b;
        i1 = x1; j1 = y1; i2 = s[k + 2]; j2 = s[l - 2];
      }
      const int kl = sk * 4 + sl, ij11 = ((si * 4 + sj) * 4 + i1) * 4 + j1;
      if (a == 1 && b == 1) r.u = __ldg(&T.g->int11[ij11 * 16 + kl]);
      else if (a == 1 && b == 2) r.u = __ldg(&T.g->int12[(ij11 * 4 + j2) * 16 + kl]);
      else if (a == 2 && b == 1) r.u = __ldg(&T.g->int12[((((sl * 4 + sk) * 4 + j1) * 4 + i2) * 4 + i1) * 16 + sj * 4 + si]);
      else r.u = __ldg(&T.g->int22[(ij11 * 16 + i2 * 4 + j2) * 16 + kl]);
    } else {
      const int X = (a == 1 || b == 1) ? 0 : ((a == 2 && b == 3) || (a == 3 && b == 2)) ? 1 : 2;
      r.cls = 4 + X;
      r.v = T.sm->ninio[a * 31 + b];
      r.u = T.sm->tm2[X][code];
    }
    return r;
  }
  // rows a >= 3 hold only long bulges (b == 0) and generic interior loops.  x + (-0.0f) == x bit for bit, which
  // lets the bulge skip the mismatch term without a branch.
  static constexpr int kFastFrom = 3;
  RNA_DEVM float fast(int a, int b, int code) const {
    const int X = (b == 1) ? 0 : (a == 3 && b == 2) ? 1 : 2;
    const float tf = (X == 0) ? tm_f0 : (X == 1) ? tm_f1 : tm_f2;
    const float u = T.sm->tm2[X][code];
    const float vv = (b == 0) ? T.sm->bulge_init[a] : T.sm->ninio[a * 31 + b];
    float mm = INSIDE ? __fadd_rn(tf, u) : __fadd_rn(u, tf);   // closing-side term first
    if (b == 0) mm = __int_as_float((int)0x80000000u);
    const float pen_var = augu_pair(INSIDE ? (code >> 2) & 3 : code >> 6, INSIDE ? code >> 6 : (code >> 2) & 3) ? T.g->augu_pen : 0.f;
    const float pen_close = INSIDE ? pen_fixed : pen_var, pen_encl = INSIDE ? pen_var : pen_fixed;
    return __fadd_rn(__fadd_rn(__fadd_rn(vv, mm), pen_close), pen_encl);
  }
  RNA_DEVM float score(const Term2& t) const {
    const float pen_var = t.aux ? T.g->augu_pen : 0.f;
    const float pen_close = INSIDE ? pen_fixed : pen_var, pen_encl = INSIDE ? pen_var : pen_fixed;
    float sc;
    if (t.cls == 0 || t.cls == 3) sc = t.u;
    else if (t.cls == 1) sc = __fadd_rn(t.v, t.u);
    else if (t.cls == 2) sc = __fadd_rn(__fadd_rn(t.v, pen_close), pen_encl);
    else {
      const float tf = (t.cls == 4) ? tm_f0 : (t.cls == 5) ? tm_f1 : tm_f2;
      const float mm = INSIDE ? __fadd_rn(tf, t.u) : __fadd_rn(t.u, tf);   // closing-side term first
      sc = __fadd_rn(__fadd_rn(__fadd_rn(t.v, mm), pen_close), pen_encl);
    }
    return sc;
  }
};

template <bool CONTRA, bool INSIDE> struct LoopOf;
template <bool INSIDE> struct LoopOf<true, INSIDE> { typedef ContraLoop<INSIDE> type; };
template <bool INSIDE> struct LoopOf<false, INSIDE> { typedef TurnerLoop<INSIDE> type; };

template <bool CONTRA, bool INSIDE, class SV>
RNA_DEV typename LoopOf<CONTRA, INSIDE>::type make_loop(const SV& v, const typename Model2<CONTRA>::View& T, int i, int j) {
  if constexpr (CONTRA) return ContraLoop<INSIDE>(T, v.RR, v.LL, i, j);
  else return TurnerLoop<INSIDE>(T, v.s, v.RR, v.LL, i, j);
}

// operand of the fold from a term's dynamic values and its static score:
// INSIDE: y = C(k,l) + score.   OUTSIDE: y = ((P(k,l) + C(i,j)) - C(k,l)) + score
template <bool INSIDE>
RNA_DEV float term_operand(float c, float pv, float Cij, float score) {
  if (INSIDE) return __fadd_rn(c, score);
  return __fadd_rn(__fsub_rn(__fadd_rn(pv, Cij), c), score);
}

// windows of the closable bit matrix that hold the partners of cell (i,j) in row a of its two-loop enumeration.
// INSIDE: partners k = i+1+a ascending, l descending from j-1 (a+b <= MAX2, k <= j-2).
// OUTSIDE: k = i-1-a descending, l ascending from j+1 (k >= 0, l <= L-1).
template <bool INSIDE, class SV>
struct Windows {
  const SV& v;
  int i, j, MAX2, amax, wpos, bcap;
  RNA_DEVM Windows(const SV& v_, int MAX2_, int i_, int j_) : v(v_), i(i_), j(j_), MAX2(MAX2_) {
    bcap = v.L - 2 - j;
    amax = INSIDE ? min(MAX2, j - i - 3) : ((bcap >= 0) ? min(MAX2, i - 1) : -1);
    wpos = INSIDE ? j - 32 : j + 1;
  }
  RNA_DEVM uint32_t operator()(int aa) const {
    if (aa > amax) return 0u;
    const int kk = INSIDE ? i + 1 + aa : i - 1 - aa;
    const uint32_t raw = get32(v.mask + kk * v.W2, wpos);
    if (INSIDE) return raw & (0xffffffffu << (31 - (MAX2 - aa)));
    return raw & (0xffffffffu >> (31 - min(MAX2 - aa, bcap)));
  }
  // the same without a branch (rows past amax read a safe row and are masked to 0): for loops whose lanes must not
  // diverge
  RNA_DEVM uint32_t flat(int aa) const {
    const bool valid = aa <= amax;
    const int ac = valid ? aa : 0;
    const int kk = valid ? (INSIDE ? i + 1 + ac : i - 1 - ac) : i;
    const uint32_t raw = get32(v.mask + kk * v.W2, wpos);
    const uint32_t m = INSIDE ? (0xffffffffu << (31 - (MAX2 - ac))) : (0xffffffffu >> (31 - min(MAX2 - ac, max(bcap, 0))));
    return valid ? (raw & m) : 0u;
  }
};

// The pipelined enumeration: sink(t2) is called once per term in the reference's order (and a few times with
// the neutral term q = -1, c = -inf while the pipeline fills and drains).  LOADS: fetch C (and log P).
template <bool INSIDE, bool LOADS, class SV, class LOOP, class SINK>
RNA_DEV void twoloop_foreach(const SV& v, const LOOP& lp, int MAX2, int i, int j, SINK& sink) {
  const int L = v.L;
  const float NEG = RNA_NEG_INF;
  const Windows<INSIDE, SV> window(v, MAX2, i, j);
  const int amax = window.amax;
  int a = -1, k = i, kcode = 0;
  uint32_t w = 0;
  uint32_t wnext = window(0);   // the window of the NEXT row is loaded one row ahead
  Term1 t1;  t1.c = NEG; t1.pv = NEG; t1.code = 0; t1.a = 0; t1.b = 0; t1.q = -1;
  Term2 t2 = lp.stage2(t1);
  int idle = 0;
  for (;;) {
    // ---- stage 2: term n+1 -------------------------------------------------------------------------------
    const Term2 t2n = lp.stage2(t1);
    // ---- stage 1: term n+2 -------------------------------------------------------------------------------
    while (w == 0 && a < amax) {
      a++;
      w = wnext;
      wnext = window(a + 1);
      k = INSIDE ? i + 1 + a : i - 1 - a;
      kcode = INSIDE ? v.LL[k] : v.RR[k] * 16;
    }
    t1.c = NEG; t1.pv = NEG; t1.code = 0; t1.a = 0; t1.b = 0; t1.q = -1;
    if (w != 0) {
      int l, b;
      if (INSIDE) { const int t = 31 - __clz(w); w &= ~(1u << t); l = j - 32 + t; b = 31 - t; }
      else { const int t = __ffs(w) - 1; w &= w - 1; l = j + 1 + t; b = t; }
      const int q = doff(l - k, L) + k;
      if (LOADS) {
        t1.c = v.C[q];
        if (!INSIDE) t1.pv = v.Pm[q];
      }
      t1.code = INSIDE ? v.RR[l] * 16 + kcode : kcode + v.LL[l];
      t1.a = a; t1.b = b; t1.q = q;
      idle = 0;
    } else {
      idle++;
    }
    // ---- stage 3: term n ---------------------------------------------------------------------------------
    sink(t2);
    t2 = t2n;
    if (idle >= 2) break;
  }
}

// the fold with scores computed on the fly (sequences whose term streams do not fit their workspace slot,
// and the HBM-resident mode for long sequences)
template <bool INSIDE, class SV, class LOOP>
RNA_DEV float twoloop_chain(const SV& v, const LOOP& lp, const float4* lut, int MAX2, int i, int j, float Cij,
                            float sum) {
  auto sink = [&](const Term2& t) { sum = lse(sum, term_operand<INSIDE>(t.c, t.pv, Cij, lp.score(t)), lut); };
  twoloop_foreach<INSIDE, true>(v, lp, MAX2, i, j, sink);
  return sum;
}

// =========================================================================================================
// Term streams.  The two-loop scores and the partner lists are pure functions of the sequence, and both
// passes walk them (the reference caches them in a 4-D hash map, mccaskill_algo.rs:320,431,589,696).  They
// are materialised ONCE per sequence, by all threads of the CTA, as two streams of (score, partner index)
// in the CTA's HBM workspace slot: closing-pair-major for the inside chains, enclosed-pair-major for the
// outside chains, every cell's terms in the reference's fold order.
//
// Layout: the closable cells of a diagonal are cut into GROUPS of 32 consecutive cells = the 32 lanes of the
// warp that will fold them.  A group's block is lane-interleaved, element (lane, n) at base + wd n + lane
// (wd = cells in the group), and padded to the longest chain of the group with the neutral element (score 0, partner cell (0,0), whose
// sums_close is -inf: the operand becomes -inf and logsumexp skips it, utils.rs:580-583).  So a warp reads
// 256 contiguous bytes per step, all its lanes run the same trip count, and the latency-critical chain is:
// one coalesced stream load (prefetched a block ahead), one gather of C (and log P), logsumexp.
// =========================================================================================================
template <bool INSIDE, class SV>
RNA_DEV uint32_t count_terms(const SV& v, int MAX2, int i, int j) {
  const Windows<INSIDE, SV> window(v, MAX2, i, j);
  uint32_t n = 0;
  for (int a = 0; a <= window.amax; a++) n += (uint32_t)__popc(window(a));
  return n;
}

// phase 1: per closable cell, the number of terms of its inside and outside chains
template <class SV>
RNA_DEV void stream_count(const SV& v, const ModelParams& P, int lane, int nl) {
  const int L = v.L, TRI = L * (L + 1) / 2;
  for (int d = 0; d < L; d++) {
    const int cnt = v.pcnt[d], od = doff(d, L);
    for (int r = (lane - od % nl + nl) % nl; r < cnt; r += nl) {   // round-robin over all cells
      const int i = v.plist[od + r], j = i + d;
      v.ccnt[od + r] = (uint16_t)count_terms<true>(v, P.MAX2, i, j);
      v.ccnt[TRI + od + r] = (uint16_t)count_terms<false>(v, P.MAX2, i, j);
    }
  }
}
// phase 2: per group, the longest chain -> gbin[G+1], gbout[G+1]
template <bool INSIDE, class SV>
RNA_DEV void stream_groupmax_pass(const SV& v, int tid, int nt) {
  const int L = v.L, TRI = L * (L + 1) / 2;
  const uint32_t* gcum = INSIDE ? v.gcumI : v.gcumO;
  const uint16_t* gstep = INSIDE ? v.gstepI : v.gstepO;
  uint32_t* gb = INSIDE ? v.gbin : v.gbout;
  const int ns = INSIDE ? num_steps_inside(v) : num_steps_outside(v);
  const uint32_t NG = gcum[max(ns, 0)];
  for (uint32_t G = tid; G < NG; G += nt) {
    const int st = gstep[G], c = (int)(G - gcum[st]);
    const StepCells sc = step_cells<INSIDE>(v, st);
    uint32_t mx = 0;
    for (int x = 32 * c; x < min(sc.cA + sc.cB, 32 * c + 32); x++) {
      int dd, r;
      step_cell(sc, x, dd, r);
      mx = max(mx, (uint32_t)v.ccnt[(INSIDE ? 0 : TRI) + doff(dd, L) + r]);
    }
    gb[G + 1] = mx;
  }
}
template <class SV>
RNA_DEV void stream_groupmax(const SV& v, int tid, int nt) {
  stream_groupmax_pass<true>(v, tid, nt);
  stream_groupmax_pass<false>(v, tid, nt);
}
// phase 3 (one thread): block offsets
template <bool INSIDE, class SV>
RNA_DEV void stream_scan_pass(const SV& v) {
  const uint32_t* gcum = INSIDE ? v.gcumI : v.gcumO;
  uint32_t* gb = INSIDE ? v.gbin : v.gbout;
  const int ns = INSIDE ? num_steps_inside(v) : num_steps_outside(v);
  uint32_t run = 0;
  gb[0] = 0;
  for (int st = 0; st < ns; st++) {
    const StepCells sc = step_cells<INSIDE>(v, st);
    const int tot = sc.cA + sc.cB, ng = (tot + 31) >> 5;
    for (int c = 0; c < ng; c++) {
      const uint32_t g = gcum[st] + c;
      run += group_width(tot, c) * gb[g + 1];
      gb[g + 1] = run;
    }
  }
}
template <class SV>
RNA_DEV void stream_scan(const SV& v) {
  stream_scan_pass<true>(v);
  stream_scan_pass<false>(v);
}
// phase 4: score every term once and write both streams.  Each lane walks the partner list of one cell (row a,
// bits of the row window); a warp takes 32 consecutive cells of the pass at a time.
template <bool CONTRA, bool INSIDE, bool FAST, class SV, class LOOP>
RNA_DEV uint32_t stream_fill_rows(const SV& v, const LOOP& lp, const Windows<INSIDE, SV>& window, int i, int j, int a0,
                                  int a1, uint2* out, uint32_t wd, uint32_t n) {
  const int L = v.L;
  if (a0 > a1) return n;
  int a = a0 - 1, k = i, kcode = 0;
  uint32_t w = 0, wnext = window(a0);
  for (;;) {
    if (w == 0) {           // next row (its window was loaded one row ahead)
      if (a >= a1) break;
      a++;
      w = wnext;
      wnext = (a + 1 <= a1) ? window(a + 1) : 0u;
      k = INSIDE ? i + 1 + a : i - 1 - a;
      kcode = INSIDE ? v.LL[k] : v.RR[k] * 16;
    }
    if (w != 0) {
      int l, b;
      if (INSIDE) { const int t = 31 - __clz(w); w &= ~(1u << t); l = j - 32 + t; b = 31 - t; }
      else { const int t = __ffs(w) - 1; w &= w - 1; l = j + 1 + t; b = t; }
      const int q = doff(l - k, L) + k;
      const int code = INSIDE ? v.RR[l] * 16 + kcode : kcode + v.LL[l];
      float sc;
      if (FAST) {
        sc = lp.fast(a, b, code);
      } else {
        Term1 t1;
        t1.c = 0.f; t1.pv = 0.f; t1.q = q; t1.code = code; t1.a = a; t1.b = b;
        sc = lp.score(lp.stage2(t1));
      }
      RNA_ST_STREAM(&out[wd * n], make_uint2((unsigned)__float_as_int(sc), (unsigned)q));
      n++;
    }
  }
  return n;
}
// The fast rows (a >= kFastFrom, the bulk of the terms) without divergence: every lane executes the same
// instruction stream each iteration -- advance to the next row by selects, score with safe operands, store under
// a predicate -- so a warp's 32 cells proceed in lock step and only the differences of their term counts idle.
template <bool CONTRA, bool INSIDE, class SV, class LOOP>
RNA_DEV uint32_t stream_fill_rows_flat(const SV& v, const LOOP& lp, const Windows<INSIDE, SV>& window, int i, int j,
                                       int a0, int a1, uint2* out, uint32_t wd, uint32_t n) {
  const int L = v.L;
  if (a0 > a1) return n;
  int a = a0;
  uint32_t w = window.flat(a0), wnext = window.flat(a0 + 1);
  while (a <= a1) {
    const bool has = w != 0;
    const int k = INSIDE ? i + 1 + a : i - 1 - a;
    const int kcode = INSIDE ? v.LL[k] : v.RR[k] * 16;
    const uint32_t ww = has ? w : 1u;
    int l, b;
    if (INSIDE) { const int t = 31 - __clz(ww); w = has ? (w & ~(1u << t)) : 0u; l = j - 32 + t; b = 31 - t; }
    else { const int t = __ffs(ww) - 1; w = has ? (w & (w - 1)) : 0u; l = j + 1 + t; b = t; }
    l = has ? l : (INSIDE ? k + 1 : min(j + 1, L - 1));   // a safe partner for idle lanes (never stored)
    b = has ? b : 0;
    const int q = doff(l - k, L) + k;
    const int code = INSIDE ? v.RR[l] * 16 + kcode : kcode + v.LL[l];
    const float sc = lp.fast(a, b, code);
    if (has) RNA_ST_STREAM(&out[wd * n], make_uint2((unsigned)__float_as_int(sc), (unsigned)q));
    n += has ? 1u : 0u;
    // next row when this one is exhausted (its window was loaded one row ahead)
    const bool adv = w == 0;
    const uint32_t w2 = window.flat(a + 2);
    w = adv ? wnext : w;
    wnext = adv ? w2 : wnext;
    a += adv ? 1 : 0;
  }
  return n;
}
template <bool CONTRA, bool INSIDE, class SV>
RNA_DEV void stream_fill_cell(const SV& v, const typename Model2<CONTRA>::View& T, const ModelParams& P, int i, int j,
                              uint2* out, uint32_t wd, uint32_t nmax) {
  typedef typename LoopOf<CONTRA, INSIDE>::type LOOP;
  const LOOP lp = make_loop<CONTRA, INSIDE>(v, T, i, j);
  const Windows<INSIDE, SV> window(v, P.MAX2, i, j);
  uint32_t n = 0;
  n = stream_fill_rows<CONTRA, INSIDE, false>(v, lp, window, i, j, 0, min(LOOP::kFastFrom - 1, window.amax), out, wd, n);
  n = stream_fill_rows_flat<CONTRA, INSIDE>(v, lp, window, i, j, LOOP::kFastFrom, window.amax, out, wd, n);
  for (; n < nmax; n++) RNA_ST_STREAM(&out[wd * n], make_uint2(0u, 0u));   // neutral padding
}
// fill tasks: chunks of 32 consecutive cells of a pass; task tau -> (pass, chunk), longest partner lists first
template <class SV>
RNA_DEV uint32_t stream_num_tasks(const SV& v) {
  const uint32_t nci = v.ccumI[max(num_steps_inside(v), 0)], nco = v.ccumO[max(num_steps_outside(v), 0)];
  return ((nci + 31) >> 5) + ((nco + 31) >> 5);
}
template <bool CONTRA, bool INSIDE, class SV>
RNA_DEV void stream_fill_chunk(const SV& v, const typename Model2<CONTRA>::View& T, const ModelParams& P, uint32_t cell) {
  const int L = v.L;
  const uint32_t* ccum = INSIDE ? v.ccumI : v.ccumO;
  const uint32_t* gcum = INSIDE ? v.gcumI : v.gcumO;
  const uint32_t* gb = INSIDE ? v.gbin : v.gbout;
  const int ns = INSIDE ? num_steps_inside(v) : num_steps_outside(v);
  if (cell >= ccum[max(ns, 0)]) return;
  int lo = 0, hi = ns - 1;            // last step with ccum[st] <= cell
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (ccum[mid] <= cell) lo = mid; else hi = mid - 1;
  }
  const int st = lo, x = (int)(cell - ccum[st]);
  const StepCells sc = step_cells<INSIDE>(v, st);
  int dd, r;
  step_cell(sc, x, dd, r);
  const int i = v.plist[doff(dd, L) + r];
  const uint32_t G = gcum[st] + (uint32_t)(x >> 5), wd = group_width(sc.cA + sc.cB, x >> 5), g0 = gb[G];
  stream_fill_cell<CONTRA, INSIDE>(v, T, P, i, i + dd, (INSIDE ? v.tin : v.tout) + g0 + (x & 31), wd, (gb[G + 1] - g0) / wd);
}
template <bool CONTRA, class SV>
RNA_DEV void stream_fill_task(const SV& v, const typename Model2<CONTRA>::View& T, const ModelParams& P, uint32_t tau,
                              int ln) {
  const uint32_t nci = v.ccumI[max(num_steps_inside(v), 0)], nco = v.ccumO[max(num_steps_outside(v), 0)];
  const uint32_t NKI = (nci + 31) >> 5, NKO = (nco + 31) >> 5;
  // interleave the two passes, last chunks (longest partner lists) first
  const uint32_t both = 2 * min(NKI, NKO);
  bool inside;
  uint32_t K;
  if (tau < both) { inside = (tau & 1u) == 0; K = (inside ? NKI : NKO) - 1 - (tau >> 1); }
  else { inside = NKI > NKO; K = (inside ? NKI : NKO) - 1 - (tau - both) - (both >> 1); }
  if (inside) stream_fill_chunk<CONTRA, true>(v, T, P, 32 * K + (uint32_t)ln);
  else stream_fill_chunk<CONTRA, false>(v, T, P, 32 * K + (uint32_t)ln);
}
// the latency-critical fold over a lane's column of its group's block.  Every step of the warp touches a new
// 256-byte line pair, so the stream is read a block of 8 steps ahead (>= 800 cycles of logsumexp), and the
// gathers of a block are issued before its chain starts.
#ifndef RNA_STREAM_BLOCK
#define RNA_STREAM_BLOCK 4
#endif
#ifndef RNA_STREAM_PF_STEP
#define RNA_STREAM_PF_STEP 2
#endif
#ifndef RNA_STREAM_PF_DIST
#define RNA_STREAM_PF_DIST 32
#endif
template <bool INSIDE, class SV>
RNA_DEV float stream_chain(const SV& v, const uint2* __restrict__ st, uint32_t wd, uint32_t n, const float4* lut,
                           float Cij, float sum) {
  constexpr int B = RNA_STREAM_BLOCK;
  if (n == 0) return sum;
  ChainSum cs = chain_begin(sum);
  uint2 nx[B];
#pragma unroll
  for (int k = 0; k < B; k++) nx[k] = ((uint32_t)k < n) ? RNA_LD_STREAM(&st[wd * k]) : make_uint2(0u, 0u);
  for (uint32_t pos = 0; pos < n; pos += B) {
    uint2 cur[B];
    float c[B], p[B];
#pragma unroll
    for (int k = 0; k < B; k++) cur[k] = nx[k];
#pragma unroll
    for (int k = 0; k < B; k++) nx[k] = (pos + B + k < n) ? RNA_LD_STREAM(&st[wd * (pos + B + k)]) : make_uint2(0u, 0u);
    // HBM latency exceeds a block of logsumexp's: pull the lines into L2 well ahead (one 8-byte element per
    // lane and step => a warp's step covers at most 256 bytes; every other step touches all lines)
#pragma unroll
    for (int k = 0; k < B; k += RNA_STREAM_PF_STEP)
      if (pos + RNA_STREAM_PF_DIST + k < n) RNA_PREFETCH_L2(st + wd * (pos + RNA_STREAM_PF_DIST + k));
#pragma unroll
    for (int k = 0; k < B; k++) {
      c[k] = v.C[cur[k].y];                    // neutral element: C[0] = -inf => operand -inf => no-op
      p[k] = INSIDE ? 0.f : v.Pm[cur[k].y];
    }
#pragma unroll
    for (int k = 0; k < B; k++)
      chain_add(cs, term_operand<INSIDE>(c[k], p[k], Cij, __int_as_float((int)cur[k].x)), lut);
  }
  return chain_end(cs);
}

// The same fold for a view whose sums_close / log P live in HBM/L2 (the cooperative kernel: CoopView): the gathers of
// a block are issued one whole block of logsumexp's (B x ~115 cycles) before they are consumed and the stream
// elements two blocks ahead, so neither the stream nor the gathers are waited for (the plain version waits for its
// gathers at the head of every block, which is fine when they are shared-memory loads).
template <bool INSIDE, int B>
RNA_DEV float stream_chain_deep(const float* __restrict__ C, const float* __restrict__ Pm, const uint2* __restrict__ st,
                                uint32_t wd, uint32_t n, const float4* lut, float Cij, float sum) {
  if (n == 0) return sum;
  uint2 e2[B];
  float c1[B], p1[B], s1[B];
#pragma unroll
  for (int k = 0; k < B; k++) {
    const uint2 e = ((uint32_t)k < n) ? RNA_LD_STREAM(&st[wd * k]) : make_uint2(0u, 0u);
    c1[k] = C[e.y]; p1[k] = INSIDE ? 0.f : Pm[e.y]; s1[k] = __int_as_float((int)e.x);
  }
#pragma unroll
  for (int k = 0; k < B; k++) e2[k] = ((uint32_t)(B + k) < n) ? RNA_LD_STREAM(&st[wd * (B + k)]) : make_uint2(0u, 0u);
#pragma unroll 1
  for (uint32_t pos = 0; pos < n; pos += B) {
    float c0[B], p0[B], s0[B];
#pragma unroll
    for (int k = 0; k < B; k++) { c0[k] = c1[k]; p0[k] = p1[k]; s0[k] = s1[k]; }
#pragma unroll
    for (int k = 0; k < B; k++) {   // gathers of the next block (neutral element 0: C[0] = -inf, the fold ignores it)
      c1[k] = C[e2[k].y]; p1[k] = INSIDE ? 0.f : Pm[e2[k].y]; s1[k] = __int_as_float((int)e2[k].x);
    }
#pragma unroll
    for (int k = 0; k < B; k++) e2[k] = (pos + 2 * B + k < n) ? RNA_LD_STREAM(&st[wd * (pos + 2 * B + k)]) : make_uint2(0u, 0u);
#pragma unroll
    for (int k = 0; k < B; k += 2)
      if (pos + 48 + k < n) RNA_PREFETCH_L2(st + wd * (pos + 48 + k));
#pragma unroll
    for (int k = 0; k < B; k++) sum = RNA_COOP_LSE(sum, term_operand<INSIDE>(c0[k], p0[k], Cij, s0[k]), lut);
  }
  return sum;
}
RNA_CHAIN_FN float coop_stream_fold(const float* C, const float* Pm, const uint2* st, uint32_t wd, uint32_t n,
                                    const float4* lut, float Cij, float sum, int inside) {
  if (inside) return stream_chain_deep<true, 8>(C, Pm, st, wd, n, lut, Cij, sum);
  return stream_chain_deep<false, 8>(C, Pm, st, wd, n, lut, Cij, sum);
}
RNA_CHAIN_PTR(coop_stream_fold)
// compile-time choice by view type
template <class SV> struct StreamDepth { static constexpr int value = 0; };   // 0: plain stream_chain
template <> struct StreamDepth<CoopView> { static constexpr int value = 8; };
template <bool INSIDE, class SV>
RNA_DEV float stream_fold(const SV& v, const uint2* __restrict__ st, uint32_t wd, uint32_t n, const float4* lut,
                          float Cij, float sum) {
  if constexpr (StreamDepth<SV>::value > 0) return RNA_CHAIN_CALL(coop_stream_fold)(v.C, v.Pm, st, wd, n, lut, Cij, sum, INSIDE ? 1 : 0);
  else return stream_chain<INSIDE>(v, st, wd, n, lut, Cij, sum);
}

// =========================================================================================================
// inside, role X: sums_close of the closable cells of diagonal d (src/mccaskill_algo.rs:290-343, 395-467).
// Needs: sums_close of diagonals <= d-2, sums_multibranch of diagonal d-2.
// =========================================================================================================
// hairpin + two-loop part of sums_close of the x-th closable cell of inside step st (no multibranch term yet)
template <bool CONTRA, class SV>
RNA_DEV float inside_cell_partial(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                                  const ModelParams& P, int st, int tot, int x, int d, int i, int j) {
  const uint8_t* s = v.s;
  float sum = RNA_NEG_INF;   // (the first fold into the empty sum just takes the finite operand)
  if constexpr (CONTRA) {
    if (d - 1 <= P.MAX2) sum = lse_init(c2_hairpin(T, s, i, j));
  } else {
    sum = lse_init(t_hairpin(T, s, i, j));
  }
  if (v.tin) {
    const uint32_t G = v.gcumI[st] + (x >> 5), gb = v.gbin[G], wd = group_width(tot, x >> 5);
    sum = stream_fold<true>(v, v.tin + gb + (x & 31), wd, (v.gbin[G + 1] - gb) / wd, lut, 0.f, sum);
  } else {
    typename LoopOf<CONTRA, true>::type lp = make_loop<CONTRA, true>(v, T, i, j);
    sum = twoloop_chain<true>(v, lp, lut, P.MAX2, i, j, 0.f, sum);
  }
  return sum;
}
// the closing multibranch term (needs sums_multibranch of diagonal d-2)
template <bool CONTRA, class SV>
RNA_DEV float inside_cell_fin(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d, int i, int j,
                              float sum) {
  const float* Mm2 = v.Mroll + ((d + 1) % 3) * v.L;   // diagonal d-2
  const float mb = (d >= 2) ? Mm2[i + 1] : RNA_NEG_INF;
  return lse(sum, __fadd_rn(mb, v2_mbclose<CONTRA>(T, v.s, v.L, i, j)), lut);
}
// X, phase 1 of inside step st: hairpin + two-loop part of sums_close for the diagonals d and d+1 of the step.
// Both only need sums_close of diagonals <= d-1, so the two longest chains of the pass run side by side.
template <bool CONTRA, class SV>
RNA_DEV void inside_X(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                      const ModelParams& P, int st, int lane, int nl) {
  const int L = v.L;
  const StepCells sc = step_cells<true>(v, st);
  const int tot = sc.cA + sc.cB;
  for (int x = lane; x < tot; x += nl) {   // ONE loop body: lanes of both diagonals run their chains together
    int d, r;
    step_cell(sc, x, d, r);
    const int od = doff(d, L), i = v.plist[od + r];
    // still without the multibranch term: no one reads these diagonals before inside_X_fin
    v.C[od + i] = inside_cell_partial<CONTRA>(v, T, lut, P, st, tot, x, d, i, i + d);
  }
}
// X, phase 2: the closing multibranch term (needs sums_multibranch of d-2 resp. d-1, computed by Z in phase 1)
template <bool CONTRA, class SV>
RNA_DEV void inside_X_fin(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int st, int lane,
                          int nl) {
  const int L = v.L;
  const StepCells sc = step_cells<true>(v, st);
  for (int x = lane; x < sc.cA + sc.cB; x += nl) {
    int dd, r;
    step_cell(sc, x, dd, r);
    const int od = doff(dd, L), i = v.plist[od + r];
    v.C[od + i] = inside_cell_fin<CONTRA>(v, T, lut, dd, i, i + dd, v.C[od + i]);
  }
}
// X of a ONE-diagonal step (the grid-wide wavefront for long sequences: a barrier costs microseconds there, so a
// step is one diagonal and the whole fold of a cell runs in it; needs sums_close <= d-2 and sums_multibranch(d-2))
template <bool CONTRA, class SV>
RNA_DEV void inside_X_diag(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                           const ModelParams& P, int d, int lane, int nl) {
  const int L = v.L;
  if (d < v.din0) return;
  const int st = (d - v.din0) >> 1;
  const StepCells sc = step_cells<true>(v, st);
  const int tot = sc.cA + sc.cB, x0 = (d == sc.dA) ? 0 : sc.cA, cnt = v.pcnt[d], od = doff(d, L);
  for (int r = lane; r < cnt; r += nl) {
    const int i = v.plist[od + r], j = i + d;
    const float sum = inside_cell_partial<CONTRA>(v, T, lut, P, st, tot, x0 + r, d, i, j);
    v.C[od + i] = inside_cell_fin<CONTRA>(v, T, lut, d, i, j, sum);
  }
}

// ascending iterator over the set bits of a bit-matrix row within [lo, hi]
struct RowBits {
  const uint32_t* row;
  int p, hi;
  uint32_t w;
  RNA_DEVM RowBits(const uint32_t* row_, int lo, int hi_) : row(row_), p(lo), hi(hi_), w(0u) {
    if (p <= hi) w = window();
  }
  RNA_DEVM uint32_t window() const {
    uint32_t x = get32(row, p);
    const int n = hi - p + 1;
    if (n < 32) x &= (1u << n) - 1u;
    return x;
  }
  RNA_DEVM int next() {   // -1 when exhausted
    if (p > hi) return -1;
    while (w == 0) {
      p += 32;
      if (p > hi) return -1;
      w = window();
    }
    const int t = __ffs(w) - 1;
    w &= w - 1;
    return p + t;
  }
};

// =========================================================================================================
// inside, role Y (CONTRAfold): the k < j part of sums_rightmost_basepairs_{external,multibranch}[i][j]
// (src/mccaskill_algo.rs:468-486); the k == j term is added by role Z one step later.
// Needs: sums_close of diagonals < d.
// =========================================================================================================
template <int CH, class SV>
RNA_DEV void inside_Y_contra(const SV& v, const ContraView2& T, const float4* lut, int d, int lane, int nl,
                                int kskip = 0) {   // kskip: leave the last kskip values of k (j-1, ...) to a later phase
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  const DevContra* dev = T.g;
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    ChainSum r = chain_begin(NEG), rm = chain_begin(NEG);
    if constexpr (CH <= 1) {
    const uint32_t* row = v.mask + i * v.W2;
    for (int p = i + 1; p <= j - 1 - kskip; p += 32) {
      uint32_t w = get32(row, p);
      const int n = j - kskip - p;
      if (n < 32) w &= (1u << n) - 1u;
      while (w) {
        const int t = __ffs(w) - 1;
        w &= w - 1;
        const int k = p + t;
        const float av = __fadd_rn(v.C[doff(k - i, L) + i], v2_acc<true>(T, s, L, i, k));
        const float nn = (float)(j - k);
        chain_add(r, __fadd_rn(__fadd_rn(av, dev->ext_bp), __fmul_rn(dev->ext_unpair, nn)), lut);
        chain_add(rm, __fadd_rn(__fadd_rn(av, dev->mb_bp), __fmul_rn(dev->mb_unpair, nn)), lut);
      }
    }
    } else {
    // closable (i,k), i < k < j ascending, CH terms at a time: the sums_close gathers of the next chunk are in
    // flight while the folds of the current one execute
    RowBits it(v.mask + i * v.W2, i + 1, j - 1 - kskip);
    int kb[CH];
    float cb[CH];
#pragma unroll
    for (int u = 0; u < CH; u++) { kb[u] = it.next(); cb[u] = (kb[u] >= 0) ? v.C[doff(kb[u] - i, L) + i] : NEG; }
    while (kb[0] >= 0) {
      int ka[CH];
      float ca[CH];
#pragma unroll
      for (int u = 0; u < CH; u++) { ka[u] = kb[u]; ca[u] = cb[u]; }
#pragma unroll
      for (int u = 0; u < CH; u++) { kb[u] = it.next(); cb[u] = (kb[u] >= 0) ? v.C[doff(kb[u] - i, L) + i] : NEG; }
#pragma unroll
      for (int u = 0; u < CH; u++) {
        if (ka[u] >= 0) {
          const int k = ka[u];
          const float av = __fadd_rn(ca[u], v2_acc<true>(T, s, L, i, k));
          const float nn = (float)(j - k);
          chain_add(r, __fadd_rn(__fadd_rn(av, dev->ext_bp), __fmul_rn(dev->ext_unpair, nn)), lut);
          chain_add(rm, __fadd_rn(__fadd_rn(av, dev->mb_bp), __fmul_rn(dev->mb_unpair, nn)), lut);
        }
      }
    }
    }
    v.R[od + i] = chain_end(r);
    v.X[od + i] = chain_end(rm);
  }
}

// =========================================================================================================
// inside, role Z: finish R (/Rm) of diagonal d with the k == j term, then the three dense chains
// sums_external, sums_multibranch, sums_1ormore_basepairs (src/mccaskill_algo.rs:344-374, 468-512).
// Needs: sums_close(d) (role X, previous step), partial R/Rm(d) (role Y, previous step), R/Rm/E/M1 of < d.
// =========================================================================================================
template <bool CONTRA, int PF, class SV, bool SUMSX = false>
RNA_DEV void inside_Z(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d,
                      int lane, int nl) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  const typename Model2<CONTRA>::Dev* dev = T.g;
  float* Mcur = v.Mroll + (d % 3) * L;
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    const float c = v.C[od + i];
    const float accv = (c > NEG) ? __fadd_rn(c, v2_acc<CONTRA>(T, s, L, i, j)) : NEG;
    float Rij, Rmij = NEG;
    if constexpr (!CONTRA) {
      // prefix property of the left-to-right fold: R[i][j] = R[i][j-1] (+) A(i,j)
      const float prev = (d >= 1) ? v.R[doff(d - 1, L) + i] : NEG;
      Rij = lse(prev, accv, lut);
    } else {
      Rij = lse(v.R[od + i], __fadd_rn(__fadd_rn(accv, dev->ext_bp), __fmul_rn(dev->ext_unpair, 0.f)), lut);
      Rmij = lse(v.X[od + i], __fadd_rn(__fadd_rn(accv, dev->mb_bp), __fmul_rn(dev->mb_unpair, 0.f)), lut);
      v.X[od + i] = Rmij;
    }
    v.R[od + i] = Rij;
    ChainSum sE, sM1, sM = chain_begin(NEG);
    if constexpr (CONTRA) {
      sE = chain_begin(__fmul_rn(dev->ext_unpair, (float)(d + 1)));
      sM1 = chain_begin(Rmij);
    } else {
      sE = chain_begin(0.f);
      sM1 = chain_begin(__fadd_rn(Rij, dev->coeff_num_branches));
    }
    chain_add(sE, __fadd_rn(Rij, 0.f), lut);   // k = i: E[i][i-1] = 0 (lower triangle / literal 0)
    if constexpr (PF <= 2) {
    // operands two split points ahead are in flight while the three folds of the current one execute
    // (R, Rm, E, M1 may live in HBM/L2: their addresses do not depend on the running sums)
    float r1 = NEG, e1 = NEG, q1 = NEG, x1 = NEG, r2 = NEG, e2 = NEG, q2 = NEG, x2 = NEG;
    if (1 < d) {
      r1 = v.R[doff(d - 1, L) + i + 1]; e1 = v.E[doff(0, L) + i]; q1 = v.M1[doff(0, L) + i];
      if (CONTRA) x1 = v.X[doff(d - 1, L) + i + 1];
    }
    if (2 < d) {
      r2 = v.R[doff(d - 2, L) + i + 2]; e2 = v.E[doff(1, L) + i]; q2 = v.M1[doff(1, L) + i];
      if (CONTRA) x2 = v.X[doff(d - 2, L) + i + 2];
    }
    for (int m = 1; m < d; m++) {
      const float r = r1, e = e1, m1 = q1, rm = x1;
      r1 = r2; e1 = e2; q1 = q2; x1 = x2;
      if (m + 2 < d) {
        r2 = v.R[doff(d - m - 2, L) + i + m + 2]; e2 = v.E[doff(m + 1, L) + i]; q2 = v.M1[doff(m + 1, L) + i];
        if (CONTRA) x2 = v.X[doff(d - m - 2, L) + i + m + 2];
      }
      chain_add(sE, __fadd_rn(r, e), lut);
      if constexpr (CONTRA) {
        chain_add(sM1, __fadd_rn(rm, __fmul_rn(dev->mb_unpair, (float)m)), lut);
        chain_add(sM, __fadd_rn(m1, rm), lut);
      } else {
        const float xx = __fadd_rn(r, dev->coeff_num_branches);
        chain_add(sM1, xx, lut);
        chain_add(sM, __fadd_rn(m1, xx), lut);
      }
    }
    } else {
    // operands PF split points ahead are in flight while the three folds of the current one execute (R, Rm, E, M1
    // may live in HBM/L2: their addresses do not depend on the running sums); PF ~ memory latency / fold latency
    float pr[PF], pe[PF], pq[PF], px[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) {
      pr[u] = NEG; pe[u] = NEG; pq[u] = NEG; px[u] = NEG;
      const int m = 1 + u;
      if (m < d) {
        pr[u] = v.R[doff(d - m, L) + i + m]; pe[u] = v.E[doff(m - 1, L) + i]; pq[u] = v.M1[doff(m - 1, L) + i];
        if (CONTRA) px[u] = v.X[doff(d - m, L) + i + m];
      }
    }
    for (int m0 = 1; m0 < d; m0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; u++) {
        const int m = m0 + u;
        if (m < d) {
          const float r = pr[u], e = pe[u], m1 = pq[u], rm = px[u];
          const int mn = m + PF;
          if (mn < d) {
            pr[u] = v.R[doff(d - mn, L) + i + mn]; pe[u] = v.E[doff(mn - 1, L) + i]; pq[u] = v.M1[doff(mn - 1, L) + i];
            if (CONTRA) px[u] = v.X[doff(d - mn, L) + i + mn];
          }
          chain_add(sE, __fadd_rn(r, e), lut);
          if constexpr (CONTRA) {
            chain_add(sM1, __fadd_rn(rm, __fmul_rn(dev->mb_unpair, (float)m)), lut);
            chain_add(sM, __fadd_rn(m1, rm), lut);
          } else {
            const float xx = __fadd_rn(r, dev->coeff_num_branches);
            chain_add(sM1, xx, lut);
            chain_add(sM, __fadd_rn(m1, xx), lut);
          }
        }
      }
    }
    }
    v.E[od + i] = chain_end(sE);
    const float Mij = chain_end(sM);
    Mcur[i] = Mij;
    if constexpr (SUMSX) { if (v.Mfull) v.Mfull[sums_index(L, i, j)] = Mij; }
    v.M1[od + i] = lse(chain_end(sM1), Mij, lut);
  }
}

// =========================================================================================================
// PAIR-STEP schedule of the cooperative long-sequence kernel (one sequence on the whole GPU).  There a lane is
// cheap and latency is everything, so the three dense chains of a cell run on three different warps (one
// logsumexp per split point each instead of three interleaved ones), and the chains of TWO diagonals run side by
// side.  Inside step s, t = din0 + 2s:
//   phase A   X: hairpin + two-loop part of sums_close(t), sums_close(t+1)                (inside_X)
//             Y: CONTRAfold rightmost-pair sums of t over k <= j-1 and of t+1 over k <= j-2 (inside_Y_contra)
//             Z: chains sums_external | sums_multibranch | sums_1ormore of t-2 and t-1     (inside_chain_pair)
//   phase B   every thread: finish sums_1ormore(t-2), (t-1) with sums_multibranch; closing multibranch term of
//             sums_close(t), (t+1); the k = j-1 / k = j terms of the rightmost-pair sums   (inside_fin_pair)
// Dependencies: the chains of diagonal d read R/Rm of diagonals <= d (R(d) itself only as their first operand),
// E and M1 of diagonals <= d-2; sums_close(d) needs sums_multibranch(d-2) only in its last term.
// =========================================================================================================
// sum = lse(sum, op(m, A[i+m][j], B[i][i+m-1])) for m = 1 .. d-1, operands PF split points ahead in flight
// (A, B live in HBM/L2; their addresses are affine in m).  A: R or Rm (column j), B: E or M1 (row i) or null.
template <int PF, bool HASB, class OP>
RNA_DEV float chain_fold(const float* __restrict__ A, const float* __restrict__ B, int L, int d, int i, float sum,
                         const float4* lut, OP op) {
  // Branch-free folds: operands beyond m = d-1 are -inf, which the fold ignores (measured on B200,
  // tools/chain_microbench.cu: 145 cycles per split point at PF = 4 against 227 with a guarded fold; the bare
  // logsumexp chain is 115).
  const float NEG = RNA_NEG_INF;
  const float* pA = A + (doff(d - 1, L) + i + 1);   // operands of split point mL = 1
  const float* pB = B + i;
  int sA = d - 1 - L, sB = L, mL = 1;               // pointer steps to split point mL + 1
  float pa[PF], pb[PF];
  auto load = [&](float& a, float& b) {
    a = NEG; b = NEG;
    if (mL < d) { a = *pA; if (HASB) b = *pB; }
    pA += sA; sA--;        // doff(d-m-1) + i+m+1  -  (doff(d-m) + i+m) = d - m - L
    if (HASB) { pB += sB; sB--; }   // doff(m) - doff(m-1) = L - m + 1
    mL++;
  };
#pragma unroll
  for (int u = 0; u < PF; u++) load(pa[u], pb[u]);
#pragma unroll 1
  for (int m0 = 1; m0 < d; m0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const float a = pa[u], b = pb[u];
      load(pa[u], pb[u]);
      sum = RNA_COOP_LSE_NN(sum, op(m0 + u, a, b), lut);
    }
  }
  return sum;
}
// ---------------------------------------------------------------------------------------------------------
// The long chains of the cooperative kernel are compiled as functions of their own and CALLED THROUGH A POINTER read
// from device memory.  Inlined into the big kernel, ptxas has two predicate registers left for the seven breakpoint
// compares of the logsumexp (the others hold long-lived flags): it then interleaves the four coefficient select
// trees level by level and re-materialises their constants per fold — 142 cycles per dependent fold instead of
// ~100.  An indirect call obeys the full ABI, so the callee gets every predicate and register.  Arguments are plain
// pointers and scalars; the host build (tests/emu) calls the same functions directly.
// ---------------------------------------------------------------------------------------------------------

// dense inside chain over split points m = 1 .. d-1; opk 0: A + B, 1: A + c1*m, 2: A + c0, 3: B + (A + c0)
RNA_CHAIN_FN float coop_chain_fold(const float* A, const float* B, int L, int d, int i, float sum, int opk, float c0,
                                   float c1, const float4* lut) {
  switch (opk) {
    case 0: return chain_fold<4, true>(A, B, L, d, i, sum, lut, [](int, float a, float b) { return __fadd_rn(a, b); });
    case 1: return chain_fold<4, false>(A, B, L, d, i, sum, lut, [c1](int m, float a, float) { return __fadd_rn(a, __fmul_rn(c1, (float)m)); });
    case 2: return chain_fold<4, false>(A, B, L, d, i, sum, lut, [c0](int, float a, float) { return __fadd_rn(a, c0); });
    default: return chain_fold<4, true>(A, B, L, d, i, sum, lut, [c0](int, float a, float b) { return __fadd_rn(b, __fadd_rn(a, c0)); });
  }
}
RNA_CHAIN_PTR(coop_chain_fold)

// one dense chain (kind 0: sums_external, 1: sums_1ormore_basepairs without its last term, 2: sums_multibranch) of
// cell (i, i+d).  Needs the finished R (/Rm) of diagonal d.  src/mccaskill_algo.rs:352-374, 487-512.
template <bool CONTRA, int PF, class SV>
RNA_DEV void inside_chain_cell(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d, int kind,
                               int i) {
  const int L = v.L, od = doff(d, L);
  const typename Model2<CONTRA>::Dev* dev = T.g;
  (void)PF;
  if (kind == 0) {
    float sE;
    if constexpr (CONTRA) sE = __fmul_rn(dev->ext_unpair, (float)(d + 1)); else sE = 0.f;
    sE = lse(sE, __fadd_rn(v.R[od + i], 0.f), lut);   // k = i: E[i][i-1] = 0
    v.E[od + i] = RNA_CHAIN_CALL(coop_chain_fold)(v.R, v.E, L, d, i, sE, 0, 0.f, 0.f, lut);          // R[k][j] + E[i][k-1]
  } else if (kind == 1) {
    if constexpr (CONTRA)   // Rm[k][j] + multibranch_score_unpair * (k - i)
      v.M1[od + i] = RNA_CHAIN_CALL(coop_chain_fold)(v.X, nullptr, L, d, i, v.X[od + i], 1, 0.f, dev->mb_unpair, lut);
    else {                  // R[k][j] + COEFF_NUM_BRANCHES
      const float cb = dev->coeff_num_branches;
      v.M1[od + i] = RNA_CHAIN_CALL(coop_chain_fold)(v.R, nullptr, L, d, i, __fadd_rn(v.R[od + i], cb), 2, cb, 0.f, lut);
    }
  } else {
    float* Mcur = v.Mroll + (d % 3) * L;
    if constexpr (CONTRA)   // M1[i][k-1] + Rm[k][j]  (IEEE addition commutes bit for bit)
      Mcur[i] = RNA_CHAIN_CALL(coop_chain_fold)(v.X, v.M1, L, d, i, RNA_NEG_INF, 0, 0.f, 0.f, lut);
    else                    // M1[i][k-1] + (R[k][j] + COEFF_NUM_BRANCHES)
      Mcur[i] = RNA_CHAIN_CALL(coop_chain_fold)(v.R, v.M1, L, d, i, RNA_NEG_INF, 3, dev->coeff_num_branches, 0.f, lut);
    if (v.Mfull) v.Mfull[sums_index(L, i, i + d)] = Mcur[i];
  }
}
// Z of a pair step: warp-sized tasks (chain kind x 32 cells) of the diagonals t-2 and t-1, the longer diagonal first;
// warp w of nw takes tasks w, w + nw, ...
template <bool CONTRA, int PF, class SV>
RNA_DEV void inside_chain_pair(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int t, int w,
                               int nw, int lane32) {
  const int L = v.L;
  const int dB = t - 1, dA = t - 2;
  const int gB = (dB >= v.din0 && dB < L) ? (L - dB + 31) >> 5 : 0;
  const int gA = (dA >= v.din0 && dA < L) ? (L - dA + 31) >> 5 : 0;
  const int ntask = 3 * (gA + gB);
  for (int tau = w; tau < ntask; tau += nw) {
    const int kind = tau % 3;
    int g = tau / 3, d = dB;
    if (g >= gB) { g -= gB; d = dA; }
    const int i = g * 32 + lane32;
    if (i < L - d) inside_chain_cell<CONTRA, PF>(v, T, lut, d, kind, i);
  }
}
// CONTRAfold, Y of a pair step, DENSE: the rightmost-pair sums of cell (i, i+d) over k = i+1 .. j-1-kskip
// (src/mccaskill_algo.rs:468-486; a non-closable (i,k) has sums_close = -inf and the fold ignores it).  kind 0:
// external, kind 1: multibranch.  The accessible scores come from a table filled once (score_table_acc, in the buffer
// the outside pass later reuses for the multibranch closing scores); all loads are coalesced and affine.
template <class SV>
RNA_DEV void score_table_acc(const SV& v, const ContraView2& T, int lane, int nl) {
  const int L = v.L;
  for (int d = 0; d < L; d++) {
    const int od = doff(d, L);
    for (int i = lane; i < L - d; i += nl)
      v.MB[od + i] = (get32(v.mask + i * v.W2, i + d) & 1u) ? v2_acc<true>(T, v.s, L, i, i + d) : 0.f;
  }
}
// fold over k = i + m, m = 1 .. n, of (sums_close + accessible score + cbp) + cun * (d - m)
RNA_CHAIN_FN float coop_in_y_dense(const float* C, const float* S, int L, int d, int i, int n, float cbp, float cun,
                                   const float4* lut) {
  constexpr int PF = 4;
  const float NEG = RNA_NEG_INF;
  const float* pC = C + (L + i);     // doff(1) + i
  const float* pS = S + (L + i);
  int sC = L - 1, mL = 1;
  float rc[PF], rs[PF];
  auto load = [&](float& c, float& sc) {
    c = NEG; sc = 0.f;
    if (mL <= n) { c = *pC; sc = *pS; }
    pC += sC; pS += sC; sC--;          // doff(m+1) - doff(m) = L - m
    mL++;
  };
#pragma unroll
  for (int u = 0; u < PF; u++) load(rc[u], rs[u]);
  float sum = NEG;
#pragma unroll 1
  for (int m0 = 1; m0 <= n; m0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const float av = __fadd_rn(rc[u], rs[u]);
      load(rc[u], rs[u]);
      sum = RNA_COOP_LSE_NN(sum, __fadd_rn(__fadd_rn(av, cbp), __fmul_rn(cun, (float)(d - (m0 + u)))), lut);
    }
  }
  return sum;
}
RNA_CHAIN_PTR(coop_in_y_dense)
template <int PF, class SV>
RNA_DEV void inside_Y_dense_cell(const SV& v, const ContraView2& T, const float4* lut, int d, int kind, int i, int kskip) {
  (void)PF;
  const DevContra* dev = T.g;
  const float cbp = kind == 0 ? dev->ext_bp : dev->mb_bp, cun = kind == 0 ? dev->ext_unpair : dev->mb_unpair;
  (kind == 0 ? v.R : v.X)[doff(d, v.L) + i] = RNA_CHAIN_CALL(coop_in_y_dense)(v.C, v.MB, v.L, d, i, d - 1 - kskip, cbp, cun, lut);
}
// warp tasks (diagonal, kind, 32 cells): diagonal t+1 without k = j-1 (sums_close(t) is not finished), diagonal t whole
template <int PF, class SV>
RNA_DEV void inside_Y_dense_pair(const SV& v, const ContraView2& T, const float4* lut, int t, int w, int nw, int lane32) {
  const int L = v.L;
  const int gB = (t + 1 < L) ? (L - t - 1 + 31) >> 5 : 0, gA = (t < L) ? (L - t + 31) >> 5 : 0;
  for (int tau = w; tau < 2 * (gA + gB); tau += nw) {
    const int kind = tau & 1;
    int g = tau >> 1, d = t + 1, kskip = 1;
    if (g >= gB) { g -= gB; d = t; kskip = 0; }
    const int i = g * 32 + lane32;
    if (i < L - d) inside_Y_dense_cell<PF>(v, T, lut, d, kind, i, kskip);
  }
}
// phase B of a pair step, lane = any thread of the grid
template <bool CONTRA, class SV>
RNA_DEV void inside_fin_pair(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int t, int lane,
                             int nl) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const typename Model2<CONTRA>::Dev* dev = T.g;
  // sums_1ormore_basepairs(d) <- partial (+) sums_multibranch(d), d = t-2, t-1 (first read by the chains of d+2)
  for (int d = t - 2; d <= t - 1; d++) {
    if (d < v.din0 || d >= L) continue;
    const float* Md = v.Mroll + (d % 3) * L;
    const int od = doff(d, L);
    for (int i = lane; i < L - d; i += nl) v.M1[od + i] = lse(v.M1[od + i], Md[i], lut);
  }
  if (t >= L) return;
  const int od = doff(t, L), od1 = doff(t + 1, L);
  for (int i = lane; i < L - t; i += nl) {
    // diagonal t
    int j = i + t;
    float c = v.C[od + i];
    if (get32(v.mask + i * v.W2, j) & 1u) { c = inside_cell_fin<CONTRA>(v, T, lut, t, i, j, c); v.C[od + i] = c; }
    const float accv = (c > NEG) ? __fadd_rn(c, v2_acc<CONTRA>(T, s, L, i, j)) : NEG;
    float Rt, Rmt = NEG;
    if constexpr (!CONTRA) {
      // prefix property of the left-to-right fold: R[i][j] = R[i][j-1] (+) A(i,j)
      Rt = lse((t >= 1) ? v.R[doff(t - 1, L) + i] : NEG, accv, lut);
    } else {
      Rt = lse(v.R[od + i], __fadd_rn(__fadd_rn(accv, dev->ext_bp), __fmul_rn(dev->ext_unpair, 0.f)), lut);
      Rmt = lse(v.X[od + i], __fadd_rn(__fadd_rn(accv, dev->mb_bp), __fmul_rn(dev->mb_unpair, 0.f)), lut);
      v.X[od + i] = Rmt;
    }
    v.R[od + i] = Rt;
    // diagonal t+1, same row
    if (t + 1 >= L || i >= L - t - 1) continue;
    j = i + t + 1;
    float c1 = v.C[od1 + i];
    if (get32(v.mask + i * v.W2, j) & 1u) { c1 = inside_cell_fin<CONTRA>(v, T, lut, t + 1, i, j, c1); v.C[od1 + i] = c1; }
    const float acc1 = (c1 > NEG) ? __fadd_rn(c1, v2_acc<CONTRA>(T, s, L, i, j)) : NEG;
    if constexpr (!CONTRA) {
      v.R[od1 + i] = lse(Rt, acc1, lut);
    } else {
      // Y left out k = j-1 (sums_close(t) was not finished): A(i,j-1) is accv, one unpaired base to its right
      float r = lse(v.R[od1 + i], __fadd_rn(__fadd_rn(accv, dev->ext_bp), __fmul_rn(dev->ext_unpair, 1.f)), lut);
      float rm = lse(v.X[od1 + i], __fadd_rn(__fadd_rn(accv, dev->mb_bp), __fmul_rn(dev->mb_unpair, 1.f)), lut);
      r = lse(r, __fadd_rn(__fadd_rn(acc1, dev->ext_bp), __fmul_rn(dev->ext_unpair, 0.f)), lut);
      rm = lse(rm, __fadd_rn(__fadd_rn(acc1, dev->mb_bp), __fmul_rn(dev->mb_unpair, 0.f)), lut);
      v.R[od1 + i] = r;
      v.X[od1 + i] = rm;
    }
  }
}

// =========================================================================================================
// FoldSums / FoldScores export (rna_fold_sums_batch; src/mccaskill_algo.rs:3-22, 213-245), after the inside pass:
// lane = any thread of the team.  `out`: RNA_SUMS_PLANES planes of L(L+1)/2 floats; the sums_multibranch plane was
// written by the chains themselves (Mfull).
// =========================================================================================================
template <bool CONTRA, class SV>
RNA_DEV void export_fold_sums(const SV& v, const typename Model2<CONTRA>::View& T, const ModelParams& P, float* out,
                              int lane, int nl) {
  const int L = v.L;
  const size_t PL = (size_t)L * (L + 1) / 2;
  const float NEG = RNA_NEG_INF;
  for (int d = 0; d < L; d++) {
    const int od = doff(d, L);
    const bool visited = d >= v.din0;   // (Turner: spans below MIN_SPAN_HAIRPIN_CLOSE are never visited: init values)
    for (int i = lane; i < L - d; i += nl) {
      const int j = i + d;
      const size_t x = sums_index(L, i, j);
      const float c = v.C[od + i];
      const bool clos = (get32(v.mask + i * v.W2, j) & 1u) != 0;
      float hp = NEG, mbc = NEG, acc = NEG, A = NEG;
      if (clos) {
        if constexpr (CONTRA) { if (d - 1 <= P.MAX2) hp = c2_hairpin(T, v.s, i, j); }
        else hp = t_hairpin(T, v.s, i, j);
        if (c > NEG) {   // (the reference inserts these three only when the pair's sum is finite)
          mbc = v2_mbclose<CONTRA>(T, v.s, L, i, j);
          acc = v2_acc<CONTRA>(T, v.s, L, i, j);
          A = __fadd_rn(c, acc);
        }
      }
      out[0 * PL + x] = c;
      out[1 * PL + x] = A;
      out[2 * PL + x] = v.E[od + i];
      out[3 * PL + x] = v.R[od + i];
      out[4 * PL + x] = CONTRA ? v.X[od + i] : NEG;
      if (!visited) out[5 * PL + x] = NEG;
      out[6 * PL + x] = v.M1[od + i];
      out[7 * PL + x] = hp;
      out[8 * PL + x] = mbc;
      out[9 * PL + x] = acc;
    }
  }
}

// =========================================================================================================
// outside, role Y: probs_multibranch / probs_multibranch2 of every cell of diagonal d
// (src/mccaskill_algo.rs:540-557, 641-661).  Needs: log P of diagonals > d.
// =========================================================================================================
template <bool CONTRA, int CH, class SV>
RNA_DEV void outside_Y(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d,
                       int lane, int nl) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    ChainSum pm = chain_begin(NEG), pm2 = chain_begin(NEG);
    if constexpr (CH <= 1) {
    const uint32_t* row = v.mask + i * v.W2;
    // closable (i,k), k > j ascending; the operands of the next term are fetched before the two folds of the
    // current one (sums_1ormore_basepairs may live in HBM/L2)
    int p = j + 1;
    uint32_t w = (p < L) ? get32(row, p) : 0u;
    auto next_k = [&]() -> int {
      while (w == 0) {
        p += 32;
        if (p >= L) return -1;
        w = get32(row, p);
      }
      const int t = __ffs(w) - 1;
      w &= w - 1;
      return p + t;
    };
    int k1 = (p < L) ? next_k() : -1;
    float c1 = NEG, pv1 = NEG, m11 = NEG;
    if (k1 >= 0) {
      const int q = doff(k1 - i, L) + i;
      c1 = v.C[q]; pv1 = v.Pm[q];
      m11 = (k1 - j >= 2) ? v.M1[doff(k1 - j - 2, L) + j + 1] : NEG;
    }
    while (k1 >= 0) {
      const int k = k1, m = k - j;
      const float c = c1, pv = pv1, m1 = m11;
      k1 = next_k();
      if (k1 >= 0) {
        const int q = doff(k1 - i, L) + i;
        c1 = v.C[q]; pv1 = v.Pm[q];
        m11 = (k1 - j >= 2) ? v.M1[doff(k1 - j - 2, L) + j + 1] : NEG;
      }
      const float x = __fsub_rn(__fadd_rn(pv, v2_mbclose<CONTRA>(T, s, L, i, k)), c);
      chain_add(pm, __fadd_rn(x, m1), lut);
      if constexpr (CONTRA) chain_add(pm2, __fadd_rn(x, __fmul_rn(T.g->mb_unpair, (float)(m - 1))), lut);
      else chain_add(pm2, x, lut);
    }
    } else {
    // closable (i,k), k > j ascending, CH terms at a time; the operands of the next chunk are in flight while
    // the folds of the current one execute (sums_1ormore_basepairs, and in the HBM modes everything, is far away)
    RowBits it(v.mask + i * v.W2, j + 1, L - 1);
    int kb[CH];
    float cb[CH], pb[CH], mb[CH];
    auto fetch = [&](int u) {
      kb[u] = it.next();
      cb[u] = NEG; pb[u] = NEG; mb[u] = NEG;
      if (kb[u] >= 0) {
        const int q = doff(kb[u] - i, L) + i;
        cb[u] = v.C[q]; pb[u] = v.Pm[q];
        if (kb[u] - j >= 2) mb[u] = v.M1[doff(kb[u] - j - 2, L) + j + 1];
      }
    };
#pragma unroll
    for (int u = 0; u < CH; u++) fetch(u);
    while (kb[0] >= 0) {
      int ka[CH];
      float ca[CH], pa[CH], ma[CH];
#pragma unroll
      for (int u = 0; u < CH; u++) { ka[u] = kb[u]; ca[u] = cb[u]; pa[u] = pb[u]; ma[u] = mb[u]; }
#pragma unroll
      for (int u = 0; u < CH; u++) fetch(u);
#pragma unroll
      for (int u = 0; u < CH; u++) {
        if (ka[u] >= 0) {
          const int k = ka[u], m = k - j;
          const float x = __fsub_rn(__fadd_rn(pa[u], v2_mbclose<CONTRA>(T, s, L, i, k)), ca[u]);
          chain_add(pm, __fadd_rn(x, ma[u]), lut);
          if constexpr (CONTRA) chain_add(pm2, __fadd_rn(x, __fmul_rn(T.g->mb_unpair, (float)(m - 1))), lut);
          else chain_add(pm2, x, lut);
        }
      }
    }
    }
    v.R[od + i] = chain_end(pm);
    v.X[od + i] = chain_end(pm2);
  }
}

// =========================================================================================================
// outside, role X: log P(i,j) of the closable cells of diagonal d (src/mccaskill_algo.rs:558-604, 662-719).
// Needs: log P, probs_multibranch, probs_multibranch2 of diagonals > d.
// =========================================================================================================
// exterior term + enclosing two-loops of log P of the x-th closable cell of outside step st
template <bool CONTRA, class SV>
RNA_DEV float outside_cell_partial(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                                   const ModelParams& P, float Z, int st, int tot, int x, int i, int j, float Cij) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const typename Model2<CONTRA>::Dev* dev = T.g;
  const float Aij = __fadd_rn(Cij, v2_acc<CONTRA>(T, s, L, i, j));
  const float El = (i < 1) ? 0.f : v.E0[i - 1];
  const float Er = (j > L - 2) ? 0.f : v.EL[j + 1];
  float sm;
  if constexpr (CONTRA) sm = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(El, Er), Aij), dev->ext_bp), Z);
  else sm = __fsub_rn(__fadd_rn(__fadd_rn(El, Aij), Er), Z);
  // enclosing two-loops: k descending from i-1, l ascending from j+1
  if (v.tin) {
    const uint32_t G = v.gcumO[st] + (x >> 5), gb = v.gbout[G], wd = group_width(tot, x >> 5);
    sm = stream_fold<false>(v, v.tout + gb + (x & 31), wd, (v.gbout[G + 1] - gb) / wd, lut, Cij, sm);
  } else {
    typename LoopOf<CONTRA, false>::type lp = make_loop<CONTRA, false>(v, T, i, j);
    sm = twoloop_chain<false>(v, lp, lut, P.MAX2, i, j, Cij, sm);
  }
  return sm;
}
// enclosing multiloops, k ascending 0..i-1 (needs probs_multibranch(2) of diagonals > j - i)
template <bool CONTRA, int PF, class SV>
RNA_DEV float outside_cell_ml(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int i, int j,
                              float Cij, float sm0) {
  const int L = v.L;
  const float NEG = RNA_NEG_INF;
  const typename Model2<CONTRA>::Dev* dev = T.g;
  const float Aij = __fadd_rn(Cij, v2_acc<CONTRA>(T, v.s, L, i, j));
  float sa;
  if constexpr (CONTRA) sa = __fadd_rn(Aij, dev->mb_bp); else sa = __fadd_rn(Aij, dev->coeff_num_branches);
  ChainSum sm = chain_begin(sm0);
  if constexpr (PF == 2) {
    // operands TWO steps ahead are in flight (explicit registers, no ring): probs_multibranch(2) live in HBM/L2 in
    // the shared-memory mode and one step of three dependent folds is shorter than that latency
    auto fetch = [&](int kk, float& fx, float& fp, float& fy) {
      fx = NEG; fp = NEG; fy = NEG;
      if (kk < i) {
        const int m = i - 1 - kk, q = doff(j - kk, L) + kk;
        fx = (m >= 1) ? v.M1[doff(m - 1, L) + kk + 1] : NEG;
        fp = v.X[q]; fy = v.R[q];
      }
    };
    float ax, ap, ay, bx, bp, by;
    fetch(0, ax, ap, ay);
    fetch(1, bx, bp, by);
    for (int kk = 0; kk < i; kk++) {
      const int m = i - 1 - kk;
      const float x1 = ax, p2 = ap, y = ay;
      ax = bx; ap = bp; ay = by;
      fetch(kk + 2, bx, bp, by);
      chain_add(sm, __fadd_rn(__fadd_rn(sa, p2), x1), lut);
      if constexpr (CONTRA) chain_add(sm, __fadd_rn(__fadd_rn(sa, y), __fmul_rn(dev->mb_unpair, (float)m)), lut);
      else chain_add(sm, __fadd_rn(sa, y), lut);
      chain_add(sm, __fadd_rn(__fadd_rn(sa, x1), y), lut);
    }
  } else if constexpr (PF <= 1) {
    // operands of step kk+1 are loaded before the three dependent logsumexp's of step kk
    float nx1 = NEG, np2 = NEG, ny = NEG;
    if (i > 0) {
      const int q = doff(j, L);
      nx1 = (i - 1 >= 1) ? v.M1[doff(i - 2, L) + 1] : NEG;
      np2 = v.X[q]; ny = v.R[q];
    }
    for (int kk = 0; kk < i; kk++) {
      const int m = i - 1 - kk;
      const float x1 = nx1, p2 = np2, y = ny;
      if (kk + 1 < i) {
        const int q = doff(j - kk - 1, L) + kk + 1;
        nx1 = (m - 1 >= 1) ? v.M1[doff(m - 2, L) + kk + 2] : NEG;
        np2 = v.X[q]; ny = v.R[q];
      }
      chain_add(sm, __fadd_rn(__fadd_rn(sa, p2), x1), lut);
      if constexpr (CONTRA) chain_add(sm, __fadd_rn(__fadd_rn(sa, y), __fmul_rn(dev->mb_unpair, (float)m)), lut);
      else chain_add(sm, __fadd_rn(sa, y), lut);
      chain_add(sm, __fadd_rn(__fadd_rn(sa, x1), y), lut);
    }
  } else {
  // operands PF steps ahead are in flight while the three dependent logsumexp's of a step execute
  float bx[PF], bp[PF], by[PF];
#pragma unroll
  for (int u = 0; u < PF; u++) {
    bx[u] = NEG; bp[u] = NEG; by[u] = NEG;
    const int kk = u;
    if (kk < i) {
      const int m = i - 1 - kk, q = doff(j - kk, L) + kk;
      bx[u] = (m >= 1) ? v.M1[doff(m - 1, L) + kk + 1] : NEG;
      bp[u] = v.X[q]; by[u] = v.R[q];
    }
  }
  for (int k0 = 0; k0 < i; k0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int kk = k0 + u;
      if (kk < i) {
        const int m = i - 1 - kk;
        const float x1 = bx[u], p2 = bp[u], y = by[u];
        const int kn = kk + PF;
        if (kn < i) {
          const int mn = i - 1 - kn, q = doff(j - kn, L) + kn;
          bx[u] = (mn >= 1) ? v.M1[doff(mn - 1, L) + kn + 1] : NEG;
          bp[u] = v.X[q]; by[u] = v.R[q];
        }
        chain_add(sm, __fadd_rn(__fadd_rn(sa, p2), x1), lut);
        if constexpr (CONTRA) chain_add(sm, __fadd_rn(__fadd_rn(sa, y), __fmul_rn(dev->mb_unpair, (float)m)), lut);
        else chain_add(sm, __fadd_rn(sa, y), lut);
        chain_add(sm, __fadd_rn(__fadd_rn(sa, x1), y), lut);
      }
    }
  }
  }
  return chain_end(sm);
}
// X, phase 1 of outside step st: exterior term + enclosing two-loops of log P for the diagonals d and d-1 of the
// step: both only need log P of diagonals >= d+1.
template <bool CONTRA, class SV>
RNA_DEV void outside_X(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                       const ModelParams& P, float Z, int st, int lane, int nl) {
  const int L = v.L;
  const StepCells sc = step_cells<false>(v, st);
  const int tot = sc.cA + sc.cB;
  for (int x = lane; x < tot; x += nl) {
    int d, r;
    step_cell(sc, x, d, r);
    const int od = doff(d, L), i = v.plist[od + r];
    const float Cij = v.C[od + i];
    if (!(Cij > RNA_NEG_INF)) continue;   // statically closable, but no structure closes it (CONTRAfold, rare)
    // exterior + two-loop part; outside_X_ml continues the fold
    v.Pm[od + i] = outside_cell_partial<CONTRA>(v, T, lut, P, Z, st, tot, x, i, i + d, Cij);
  }
}
// X, phase 2: enclosing multiloops (needs probs_multibranch(2) of diagonals >= d resp. d+1)
template <bool CONTRA, int PF, class SV>
RNA_DEV void outside_X_ml(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int st, int lane,
                          int nl) {
  const int L = v.L;
  const StepCells sc = step_cells<false>(v, st);
  for (int x = lane; x < sc.cA + sc.cB; x += nl) {
    int dd, r;
    step_cell(sc, x, dd, r);
    const int od = doff(dd, L), i = v.plist[od + r];
    const float Cij = v.C[od + i];
    if (!(Cij > RNA_NEG_INF)) continue;
    v.Pm[od + i] = outside_cell_ml<CONTRA, PF>(v, T, lut, i, i + dd, Cij, v.Pm[od + i]);
  }
}
// X of a ONE-diagonal step (grid-wide wavefront): the whole fold of log P(i,j); needs log P, probs_multibranch(2)
// of diagonals > d
template <bool CONTRA, int PF, class SV>
RNA_DEV void outside_X_diag(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                            const ModelParams& P, float Z, int d, int lane, int nl) {
  const int L = v.L;
  if (d < v.dout0) return;
  const int st = (L - 1 - d) >> 1;
  const StepCells sc = step_cells<false>(v, st);
  const int tot = sc.cA + sc.cB, x0 = (d == sc.dA) ? 0 : sc.cA, cnt = v.pcnt[d], od = doff(d, L);
  for (int r = lane; r < cnt; r += nl) {
    const int i = v.plist[od + r], j = i + d;
    const float Cij = v.C[od + i];
    if (!(Cij > RNA_NEG_INF)) continue;
    const float sm = outside_cell_partial<CONTRA>(v, T, lut, P, Z, st, tot, x0 + r, i, j, Cij);
    v.Pm[od + i] = outside_cell_ml<CONTRA, PF>(v, T, lut, i, j, Cij, sm);
  }
}


// =========================================================================================================
// Outside pass of the cooperative long-sequence kernel.  One diagonal per step (log P(d) needs
// probs_multibranch(d+1), which needs log P(d+2)), but every chain is laid out for latency:
//   * probs_multibranch / probs_multibranch2 run DENSE over k = j+1 .. L-1 (a non-closable (i,k) has
//     sums_close = log P = -inf, its operand is NaN -> -inf, the fold ignores it, src/utils.rs:580-583): all
//     lanes of a warp walk the same offsets, every load is coalesced and affine, the closing scores come from
//     a table (MB) filled once, and the two sums run on two different lanes;
//   * both are stored ROW-major, and sums_1ormore_basepairs is copied row-major once, so that the multiloop
//     part of log P (fixed k, lanes = neighbouring cells of a diagonal) reads neighbouring addresses.
// =========================================================================================================
// once, between the passes: lane = any thread of the grid
template <bool CONTRA, class SV>
RNA_DEV void outside_prep(const SV& v, const typename Model2<CONTRA>::View& T, int lane, int nl) {
  const int L = v.L;
  for (int d = 0; d < L; d++) {
    const int od = doff(d, L);
    for (int i = lane; i < L - d; i += nl) {
      const int j = i + d;
      v.M1rm[doff(i, L) + d] = v.M1[od + i];
      v.MB[od + i] = (get32(v.mask + i * v.W2, j) & 1u) ? v2_mbclose<CONTRA>(T, v.s, L, i, j) : 0.f;
    }
  }
}
// kind 0: probs_multibranch[i][j], kind 1: probs_multibranch2[i][j]   (src/mccaskill_algo.rs:540-557, 641-661)
// opk 0: probs_multibranch (x + sums_1ormore[j+1][k-1]), 1: probs_multibranch2 CONTRAfold (x + unp * (m-1)), 2: Turner (x)
// with x = (log P(i,k) + closing score(i,k)) - sums_close(i,k), k = j + m, m = 1 .. L-1-j
RNA_CHAIN_FN float coop_out_y_dense(const float* Pm, const float* C, const float* S, const float* M1, int L, int d, int i,
                                    int opk, float unp, const float4* lut) {
  constexpr int PF = 4;
  const int j = i + d, n = L - 1 - j;
  const float NEG = RNA_NEG_INF;
  const float* pP = Pm + (doff(d + 1, L) + i);   // (i, j+1): same offset in Pm, C, S
  const float* pC = C + (doff(d + 1, L) + i);
  const float* pS = S + (doff(d + 1, L) + i);
  const float* pM = M1 + (j + 1);                // M1[j+1][k-1] for m = 2: diagonal 0
  int sP = L - d - 1, sM = L, mL = 1;
  float rp[PF], rc[PF], rs[PF], rm[PF];
  auto load = [&](float& a, float& c, float& sc, float& m1) {
    a = NEG; c = NEG; sc = 0.f; m1 = NEG;
    if (mL <= n) {
      a = *pP; c = *pC; sc = *pS;
      if (opk == 0 && mL >= 2) m1 = *pM;
    }
    pP += sP; pC += sP; pS += sP; sP--;          // doff(d+m+1) - doff(d+m) = L - (d+m)
    if (mL >= 2) { pM += sM; sM--; }             // doff(m-1) - doff(m-2) = L - (m-2)
    mL++;
  };
#pragma unroll
  for (int u = 0; u < PF; u++) load(rp[u], rc[u], rs[u], rm[u]);
  float sum = NEG;
#pragma unroll 1
  for (int m0 = 1; m0 <= n; m0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const float a = rp[u], c = rc[u], sc = rs[u], m1 = rm[u];
      load(rp[u], rc[u], rs[u], rm[u]);
      const float x = __fsub_rn(__fadd_rn(a, sc), c);
      float y;
      if (opk == 0) y = __fadd_rn(x, m1);
      else if (opk == 1) y = __fadd_rn(x, __fmul_rn(unp, (float)(m0 + u - 1)));
      else y = x;
      sum = RNA_COOP_LSE(sum, y, lut);   // (x is NaN for a non-closable (i,k): the normalising fold)
    }
  }
  return sum;
}
RNA_CHAIN_PTR(coop_out_y_dense)
// kind 0: probs_multibranch[i][j], kind 1: probs_multibranch2[i][j]   (src/mccaskill_algo.rs:540-557, 641-661)
template <bool CONTRA, int PF, class SV>
RNA_DEV void outside_Y_dense_cell(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d,
                                  int kind, int i) {
  (void)PF;
  float unp = 0.f;
  if constexpr (CONTRA) unp = T.g->mb_unpair;
  const int opk = kind == 0 ? 0 : (CONTRA ? 1 : 2);
  (kind == 0 ? v.R : v.X)[doff(i, v.L) + d] = RNA_CHAIN_CALL(coop_out_y_dense)(v.Pm, v.C, v.MB, v.M1, v.L, d, i, opk, unp, lut);   // row-major
}
template <bool CONTRA, int PF, class SV>
RNA_DEV void outside_Y_dense(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d, int w, int nw,
                             int lane32) {
  const int L = v.L, ng = (L - d + 31) >> 5;
  for (int tau = w; tau < 2 * ng; tau += nw) {
    const int kind = tau & 1, i = (tau >> 1) * 32 + lane32;
    if (i < L - d) outside_Y_dense_cell<CONTRA, PF>(v, T, lut, d, kind, i);
  }
}
// the three operands of split point k of the multiloop part of log P(i,j) (src/mccaskill_algo.rs:596-600, 703-713):
// sa = A(i,j) + branch score, p2 = probs_multibranch2[k][j], y = probs_multibranch[k][j], x1 = sums_1ormore[k+1][i-1]
template <bool CONTRA>
RNA_DEV void ml_operands(float sa, float unp, int m, float p2, float y, float x1, float& a, float& b, float& c) {
  a = __fadd_rn(__fadd_rn(sa, p2), x1);
  if (CONTRA) b = __fadd_rn(__fadd_rn(sa, y), __fmul_rn(unp, (float)m)); else b = __fadd_rn(sa, y);
  c = __fadd_rn(__fadd_rn(sa, x1), y);
}
// enclosing multiloops of log P(i,j) from the row-major matrices, k ascending 0 .. i-1
// (src/mccaskill_algo.rs:594-601, 701-714)
template <bool CONTRA, int PF, class SV>
RNA_DEV float outside_cell_ml_rm(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int i, int j,
                                 float Cij, float sm) {
  const int L = v.L;
  const float NEG = RNA_NEG_INF;
  const typename Model2<CONTRA>::Dev* dev = T.g;
  const float Aij = __fadd_rn(Cij, v2_acc<CONTRA>(T, v.s, L, i, j));
  float sa;
  float unp = 0.f;
  if constexpr (CONTRA) { sa = __fadd_rn(Aij, dev->mb_bp); unp = dev->mb_unpair; } else sa = __fadd_rn(Aij, dev->coeff_num_branches);
  const float* pX = v.X + j;                 // probs_multibranch2[k][j], row k = 0
  const float* pR = v.R + j;                 // probs_multibranch[k][j]
  const float* pM = v.M1rm + (L + i - 2);    // sums_1ormore[k+1][i-1], row 1: doff(1) + (i-1) - 1
  int sQ = L - 1, sM = L - 2, kL = 0;
  float bx[PF], bp[PF], by[PF];
  auto load = [&](float& fx, float& fp, float& fy) {
    fx = NEG; fp = NEG; fy = NEG;
    if (kL < i) {
      fp = *pX; fy = *pR;
      if (kL < i - 1) fx = *pM;              // span i-1-k >= 1
    }
    pX += sQ; pR += sQ; sQ--;                // doff(k+1) + j-k-1 - (doff(k) + j-k) = L - k - 1
    pM += sM; sM--;                          // doff(k+2) + i-3-k - (doff(k+1) + i-2-k) = L - k - 2
    kL++;
  };
#pragma unroll
  for (int u = 0; u < PF; u++) load(bx[u], bp[u], by[u]);
  for (int k0 = 0; k0 < i; k0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const float x1 = bx[u], p2 = bp[u], y = by[u];
      load(bx[u], bp[u], by[u]);
      float oa, ob, oc;
      ml_operands<CONTRA>(sa, unp, i - 1 - (k0 + u), p2, y, x1, oa, ob, oc);
      sm = RNA_COOP_LSE_NN(sm, oa, lut);
      sm = RNA_COOP_LSE_NN(sm, ob, lut);
      sm = RNA_COOP_LSE_NN(sm, oc, lut);
    }
  }
  return sm;
}
template <bool CONTRA, int PF, class SV>
RNA_DEV void outside_X_diag_rm(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                               const ModelParams& P, float Z, int d, int lane, int nl, long long* tsplit = nullptr) {
  const int L = v.L;
  if (d < v.dout0) return;
  const int st = (L - 1 - d) >> 1;
  const StepCells sc = step_cells<false>(v, st);
  const int tot = sc.cA + sc.cB, x0 = (d == sc.dA) ? 0 : sc.cA, cnt = v.pcnt[d], od = doff(d, L);
  for (int r = lane; r < cnt; r += nl) {
    const int i = v.plist[od + r], j = i + d;
    const float Cij = v.C[od + i];
    if (!(Cij > RNA_NEG_INF)) continue;
#ifdef __CUDACC__
    const long long c0 = tsplit ? clock64() : 0;
#endif
    const float sm = outside_cell_partial<CONTRA>(v, T, lut, P, Z, st, tot, x0 + r, i, j, Cij);
#ifdef __CUDACC__
    if (tsplit) { tsplit[0] += clock64() - c0; tsplit[1] = i; }
#endif
    v.Pm[od + i] = outside_cell_ml_rm<CONTRA, PF>(v, T, lut, i, j, Cij, sm);
  }
}


// =========================================================================================================
// HBM-resident mode (fold_kernel2<.., MODE_GLOBAL>): the same chains with their operands in flight through per-lane
// cp.async rings in shared memory.  Separate functions on purpose: the shared-memory-mode kernel sits at its
// 64-register cap and its code generation reacts to anything that touches the functions it instantiates (a dead
// `if (ring)` branch in them cost it 3 % on the default bench), so that kernel keeps the functions above untouched.
// =========================================================================================================
#ifdef __CUDACC__
#ifndef RNA_Z_RING
#define RNA_Z_RING 8      // split points / terms in flight per lane (power of two)
#endif
template <bool CONTRA, class SV, bool SUMSX = false>
RNA_DEV void inside_Z_ring(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d,
                      int lane, int nl, float4* ring) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  const typename Model2<CONTRA>::Dev* dev = T.g;
  float* Mcur = v.Mroll + (d % 3) * L;
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    const float c = v.C[od + i];
    const float accv = (c > NEG) ? __fadd_rn(c, v2_acc<CONTRA>(T, s, L, i, j)) : NEG;
    float Rij, Rmij = NEG;
    if constexpr (!CONTRA) {
      // prefix property of the left-to-right fold: R[i][j] = R[i][j-1] (+) A(i,j)
      const float prev = (d >= 1) ? v.R[doff(d - 1, L) + i] : NEG;
      Rij = lse(prev, accv, lut);
    } else {
      Rij = lse(v.R[od + i], __fadd_rn(__fadd_rn(accv, dev->ext_bp), __fmul_rn(dev->ext_unpair, 0.f)), lut);
      Rmij = lse(v.X[od + i], __fadd_rn(__fadd_rn(accv, dev->mb_bp), __fmul_rn(dev->mb_unpair, 0.f)), lut);
      v.X[od + i] = Rmij;
    }
    v.R[od + i] = Rij;
    ChainSum sE, sM1, sM = chain_begin(NEG);
    if constexpr (CONTRA) {
      sE = chain_begin(__fmul_rn(dev->ext_unpair, (float)(d + 1)));
      sM1 = chain_begin(Rmij);
    } else {
      sE = chain_begin(0.f);
      sM1 = chain_begin(__fadd_rn(Rij, dev->coeff_num_branches));
    }
    chain_add(sE, __fadd_rn(Rij, 0.f), lut);   // k = i: E[i][i-1] = 0 (lower triangle / literal 0)
    {
    // HBM-resident mode: the operands of RNA_Z_RING split points are in flight per lane as 4-byte asynchronous copies
    // into a shared-memory ring of the lane's own (slot-major: conflict-free float4 rows) — DRAM / L2 latency is a few
    // thousand cycles under load there, far more than four split points of folds, and no register is held meanwhile
    constexpr int D = RNA_Z_RING;
    float4* my = ring + lane;
    // split points are issued in ascending m: the matrix offsets advance incrementally (diagonal-major layout:
    // off(d - m) + i + m moves by d - m - L per step, off(m - 1) + i by L - m + 1), the ring slot is a 32-bit shared address
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(my);
    int oR = doff(d - 1, L) + i + 1, oE = doff(0, L) + i, mi = 1;   // offsets of the next split point to issue (mi)
    auto issue = [&](int slot) {
      const unsigned dst = sbase + (unsigned)(slot * nl) * 16u;
      RNA_CP_ASYNC4_S(dst + 0u, v.R + oR);
      RNA_CP_ASYNC4_S(dst + 4u, v.E + oE);
      RNA_CP_ASYNC4_S(dst + 8u, v.M1 + oE);
      if (CONTRA) RNA_CP_ASYNC4_S(dst + 12u, v.X + oR);
      oR += d - mi - L;
      oE += L - mi + 1;
      mi++;
    };
#pragma unroll 1
    for (int sl = 0; sl < D; sl++) { if (1 + sl < d) issue(sl); RNA_CP_COMMIT(); }
#pragma unroll 1
    for (int m = 1; m < d; m++) {
      RNA_CP_WAIT(RNA_Z_RING - 1);
      const int slot = (m - 1) & (D - 1);
      const float4 o = my[slot * nl];
      if (m + D < d) issue(slot);
      RNA_CP_COMMIT();
      const float r = o.x, e = o.y, m1 = o.z, rm = o.w;
      chain_add(sE, __fadd_rn(r, e), lut);
      if constexpr (CONTRA) {
        chain_add(sM1, __fadd_rn(rm, __fmul_rn(dev->mb_unpair, (float)m)), lut);
        chain_add(sM, __fadd_rn(m1, rm), lut);
      } else {
        const float xx = __fadd_rn(r, dev->coeff_num_branches);
        chain_add(sM1, xx, lut);
        chain_add(sM, __fadd_rn(m1, xx), lut);
      }
    }
    RNA_CP_WAIT(0);
    }
    v.E[od + i] = chain_end(sE);
    const float Mij = chain_end(sM);
    Mcur[i] = Mij;
    if constexpr (SUMSX) { if (v.Mfull) v.Mfull[sums_index(L, i, j)] = Mij; }
    v.M1[od + i] = lse(chain_end(sM1), Mij, lut);
  }
}

template <class SV>
RNA_DEV void inside_Y_contra_ring(const SV& v, const ContraView2& T, const float4* lut, int d, int lane, int nl,
                                int kskip = 0, float4* ring = nullptr) {   // kskip: leave the last kskip values of k (j-1, ...) to a later phase
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  const DevContra* dev = T.g;
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    ChainSum r = chain_begin(NEG), rm = chain_begin(NEG);
    {
    // HBM-resident mode: the sums_close gathers of RNA_Z_RING terms are in flight in the lane's shared-memory ring
    // ({value, k}: the bit-matrix iterator runs that many terms ahead of the folds)
    constexpr int D = RNA_Z_RING;
    float4* my = ring + lane;
    RowBits it(v.mask + i * v.W2, i + 1, j - 1 - kskip);
    auto issue = [&](int slot) -> bool {
      const int k = it.next();
      if (k < 0) return false;
      float* dst = reinterpret_cast<float*>(my + slot * nl);
      RNA_CP_ASYNC4(dst, &v.C[doff(k - i, L) + i]);
      reinterpret_cast<int*>(dst)[1] = k;
      return true;
    };
    int pending = 0;
#pragma unroll 1
    for (int sl = 0; sl < D; sl++) { if (issue(sl)) pending++; RNA_CP_COMMIT(); }
#pragma unroll 1
    for (int n = 0; pending > 0; n++) {
      RNA_CP_WAIT(RNA_Z_RING - 1);
      const int slot = n & (D - 1);
      const float4 o = my[slot * nl];
      pending--;
      if (issue(slot)) pending++;
      RNA_CP_COMMIT();
      const int k = __float_as_int(o.y);
      const float av = __fadd_rn(o.x, v2_acc<true>(T, s, L, i, k));
      const float nn = (float)(j - k);
      chain_add(r, __fadd_rn(__fadd_rn(av, dev->ext_bp), __fmul_rn(dev->ext_unpair, nn)), lut);
      chain_add(rm, __fadd_rn(__fadd_rn(av, dev->mb_bp), __fmul_rn(dev->mb_unpair, nn)), lut);
    }
    RNA_CP_WAIT(0);
    }
    v.R[od + i] = chain_end(r);
    v.X[od + i] = chain_end(rm);
  }
}

template <bool CONTRA, class SV>
RNA_DEV void outside_Y_ring(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d,
                       int lane, int nl, float4* ring = nullptr) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    ChainSum pm = chain_begin(NEG), pm2 = chain_begin(NEG);
    {
    // HBM-resident mode: {sums_close, log P, sums_1ormore} of RNA_Z_RING terms in flight in the lane's ring, k alongside
    constexpr int D = RNA_Z_RING;
    float4* my = ring + lane;
    RowBits it(v.mask + i * v.W2, j + 1, L - 1);
    auto issue = [&](int slot) -> bool {
      const int k = it.next();
      if (k < 0) return false;
      float* dst = reinterpret_cast<float*>(my + slot * nl);
      const int q = doff(k - i, L) + i;
      RNA_CP_ASYNC4(dst + 0, &v.C[q]);
      RNA_CP_ASYNC4(dst + 1, &v.Pm[q]);
      if (k - j >= 2) RNA_CP_ASYNC4(dst + 2, &v.M1[doff(k - j - 2, L) + j + 1]);
      reinterpret_cast<int*>(dst)[3] = k;
      return true;
    };
    int pending = 0;
#pragma unroll 1
    for (int sl = 0; sl < D; sl++) { if (issue(sl)) pending++; RNA_CP_COMMIT(); }
#pragma unroll 1
    for (int n = 0; pending > 0; n++) {
      RNA_CP_WAIT(RNA_Z_RING - 1);
      const int slot = n & (D - 1);
      const float4 o = my[slot * nl];
      pending--;
      if (issue(slot)) pending++;
      RNA_CP_COMMIT();
      const int k = __float_as_int(o.w), m = k - j;
      const float m1 = (m >= 2) ? o.z : NEG;
      const float x = __fsub_rn(__fadd_rn(o.y, v2_mbclose<CONTRA>(T, s, L, i, k)), o.x);
      chain_add(pm, __fadd_rn(x, m1), lut);
      if constexpr (CONTRA) chain_add(pm2, __fadd_rn(x, __fmul_rn(T.g->mb_unpair, (float)(m - 1))), lut);
      else chain_add(pm2, x, lut);
    }
    RNA_CP_WAIT(0);
    }
    v.R[od + i] = chain_end(pm);
    v.X[od + i] = chain_end(pm2);
  }
}

template <bool CONTRA, class SV>
RNA_DEV float outside_cell_ml_ring(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int i, int j,
                              float Cij, float sm0, float4* ring = nullptr, int rstride = 0) {
  const int L = v.L;
  const float NEG = RNA_NEG_INF;
  const typename Model2<CONTRA>::Dev* dev = T.g;
  const float Aij = __fadd_rn(Cij, v2_acc<CONTRA>(T, v.s, L, i, j));
  float sa;
  if constexpr (CONTRA) sa = __fadd_rn(Aij, dev->mb_bp); else sa = __fadd_rn(Aij, dev->coeff_num_branches);
  ChainSum sm = chain_begin(sm0);
  {
    // HBM-resident mode: operands of RNA_Z_RING steps in flight in the lane's shared-memory ring (see inside_Z)
    constexpr int D = RNA_Z_RING;
    // steps are issued in ascending k: q = off(j - k) + k advances by j - k - L, off(i - 2 - k) + k + 1 by i - k - 2 - L
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(ring);
    int oq = doff(j, L), oM = (i >= 2) ? doff(i - 2, L) + 1 : 0, ki = 0;   // offsets of the next step to issue (ki)
    auto issue = [&](int slot) {
      const unsigned dst = sbase + (unsigned)(slot * rstride) * 16u;
      if (i - 1 - ki >= 1) RNA_CP_ASYNC4_S(dst + 0u, v.M1 + oM);
      RNA_CP_ASYNC4_S(dst + 4u, v.X + oq);
      RNA_CP_ASYNC4_S(dst + 8u, v.R + oq);
      oq += j - ki - L;
      oM += i - ki - 2 - L;
      ki++;
    };
#pragma unroll 1
    for (int sl = 0; sl < D; sl++) { if (sl < i) issue(sl); RNA_CP_COMMIT(); }
#pragma unroll 1
    for (int kk = 0; kk < i; kk++) {
      RNA_CP_WAIT(RNA_Z_RING - 1);
      const int slot = kk & (D - 1), m = i - 1 - kk;
      const float4 o = ring[slot * rstride];
      if (kk + D < i) issue(slot);
      RNA_CP_COMMIT();
      const float x1 = (m >= 1) ? o.x : NEG, p2 = o.y, y = o.z;
      chain_add(sm, __fadd_rn(__fadd_rn(sa, p2), x1), lut);
      if constexpr (CONTRA) chain_add(sm, __fadd_rn(__fadd_rn(sa, y), __fmul_rn(dev->mb_unpair, (float)m)), lut);
      else chain_add(sm, __fadd_rn(sa, y), lut);
      chain_add(sm, __fadd_rn(__fadd_rn(sa, x1), y), lut);
    }
    RNA_CP_WAIT(0);
  }
  return chain_end(sm);
}

template <bool CONTRA, class SV>
RNA_DEV void outside_X_ml_ring(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int st, int lane,
                          int nl, float4* ring) {
  const int L = v.L;
  const StepCells sc = step_cells<false>(v, st);
  for (int x = lane; x < sc.cA + sc.cB; x += nl) {
    int dd, r;
    step_cell(sc, x, dd, r);
    const int od = doff(dd, L), i = v.plist[od + r];
    const float Cij = v.C[od + i];
    if (!(Cij > RNA_NEG_INF)) continue;
    v.Pm[od + i] = outside_cell_ml_ring<CONTRA, SV>(v, T, lut, i, i + dd, Cij, v.Pm[od + i], ring ? ring + lane : nullptr, nl);
  }
}
#endif   // __CUDACC__

}  // namespace rna
