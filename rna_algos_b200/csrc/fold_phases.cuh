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
  float* C;               // sums_close
  float* R;               // sums_rightmost_basepairs_external   -> outside: probs_multibranch
  float* X;               // CONTRAfold: sums_rightmost_basepairs_multibranch -> outside: probs_multibranch2
  float* E;               // sums_external                        -> outside: log P(i,j) -> BPP
  float* M1;              // sums_1ormore_basepairs
  float* Mroll;           // sums_multibranch, 3 rolling diagonals of L
  float* E0;              // sums_external[0][x]
  float* EL;              // sums_external[x][L-1]
};

RNA_DEV int doff(int d, int L) { return d * L - ((d * (d - 1)) >> 1); }

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

// ---- CONTRAfold scorers over ContraSmall2 (bit-identical to c_* in scorers.cuh, fewer loads) ------------------
struct ContraView2 {
  const DevContra* g;
  const ContraSmall2* sm;
};
RNA_DEV float c2_js(const ContraView2& T, const uint8_t* s, int p0, int p1) {
  return T.sm->js[idx4(s[p0], s[p1], s[p0 + 1], s[p1 - 1])];
}
RNA_DEV float c2_junction(const ContraView2& T, const uint8_t* s, int L, int p0, int p1) {
  const int x = s[p0], y = s[p1];
  float v = __fadd_rn(T.sm->hc[x * 4 + y], (p0 < L - 1) ? T.sm->dl[idx3(x, y, s[min(p0 + 1, L - 1)])] : 0.f);
  return __fadd_rn(v, (p1 > 0) ? T.sm->dr[idx3(x, y, s[max(p1 - 1, 0)])] : 0.f);
}
RNA_DEV float c2_hairpin(const ContraView2& T, const uint8_t* s, int i, int j) {
  return __fadd_rn(T.g->hairpin_cum[min(j - i - 1, T.g->max_loop_len)], c2_js(T, s, i, j));
}
struct C2Outer { float js; int pq, p1, q1; };   // js(i,j); s[i]*4+s[j]; s[i+1]; s[j-1]
struct C2Inner { float js, bp; int kl; };       // js(j,i); basepair_scores[s[i]][s[j]]; s[i]*4+s[j]

// get_2loop_score_contra (src/utils.rs:423-520) with the closing pair fixed: (k,l) enclosed, a = k-i-1, b = j-l-1
RNA_DEV float c2_twoloop_outer(const ContraView2& T, const uint8_t* s, const C2Outer& o, int k, int l, int a, int b) {
  const int sk = s[k], sl = s[l];
  float sc;
  if ((a | b) == 0) {
    sc = T.sm->stack[o.pq * 16 + sk * 4 + sl];
  } else {
    float pv;
    if (a + b == 1) pv = T.sm->b1[a == 1 ? o.p1 : o.q1];
    else if (a == 1 && b == 1) pv = T.sm->i11[o.p1 * 4 + o.q1];
    else pv = __ldg(&T.g->ptab[a * 31 + b]);
    sc = __fadd_rn(__fadd_rn(pv, o.js), T.sm->js[idx4(sl, sk, s[l + 1], s[k - 1])]);
  }
  return __fadd_rn(sc, T.sm->bp[sk * 4 + sl]);
}
// the same with the ENCLOSED pair fixed: (p,q) closes, a = i-p-1, b = q-j-1
RNA_DEV float c2_twoloop_inner(const ContraView2& T, const uint8_t* s, const C2Inner& in, int p, int q, int a, int b) {
  const int sp = s[p], sq = s[q];
  float sc;
  if ((a | b) == 0) {
    sc = T.sm->stack[(sp * 4 + sq) * 16 + in.kl];
  } else {
    const int p1 = s[p + 1], q1 = s[q - 1];
    float pv;
    if (a + b == 1) pv = T.sm->b1[a == 1 ? p1 : q1];
    else if (a == 1 && b == 1) pv = T.sm->i11[p1 * 4 + q1];
    else pv = __ldg(&T.g->ptab[a * 31 + b]);
    sc = __fadd_rn(__fadd_rn(pv, T.sm->js[idx4(sp, sq, p1, q1)]), in.js);
  }
  return __fadd_rn(sc, in.bp);
}

template <bool CONTRA> struct Model2;
template <> struct Model2<false> {
  typedef DevTurner Dev; typedef TurnerSmall Small; typedef TurnerView View;
};
template <> struct Model2<true> {
  typedef DevContra Dev; typedef ContraSmall2 Small; typedef ContraView2 View;
};
template <bool CONTRA>
RNA_DEV const typename Model2<CONTRA>::Small* dev_small(const typename Model2<CONTRA>::Dev* d) {
  if constexpr (CONTRA) return &d->small2; else return &d->small;
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
// inside, role X: sums_close of the closable cells of diagonal d (src/mccaskill_algo.rs:290-343, 395-467).
// Needs: sums_close of diagonals <= d-2, sums_multibranch of diagonal d-2.
// =========================================================================================================
template <bool CONTRA, class SV>
RNA_DEV void inside_X(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                      const ModelParams& P, int d, int lane, int nl) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int cnt = v.pcnt[d], od = doff(d, L);
  const float* Mm2 = v.Mroll + ((d + 1) % 3) * L;   // diagonal d-2
  for (int r = lane; r < cnt; r += nl) {
    const int i = v.plist[od + r], j = i + d;
    float sum = NEG;
    if constexpr (CONTRA) {
      if (d - 1 <= P.MAX2) sum = lse(sum, c2_hairpin(T, s, i, j), lut);
    } else {
      sum = lse(sum, t_hairpin(T, s, i, j), lut);
    }
    C2Outer o;
    if constexpr (CONTRA) { o.js = c2_js(T, s, i, j); o.pq = s[i] * 4 + s[j]; o.p1 = s[i + 1]; o.q1 = s[j - 1]; }
    // enclosed pairs: k ascending from i+1, l descending from j-1, a + b <= MAX2
    const int amax = min(P.MAX2, d - 3);
    int a = -1, k = i;
    uint32_t w = 0;
    for (;;) {
      while (w == 0 && a < amax) {
        a++;
        k = i + 1 + a;
        // window = positions [j-32, j-1] of row k; keep l >= j-1-(MAX2-a)
        w = get32(v.mask + k * v.W2, j - 32) & (0xffffffffu << (31 - (P.MAX2 - a)));
      }
      if (w == 0) break;
      const int t = 31 - __clz(w);
      w &= ~(1u << t);
      const int l = j - 32 + t, b = 31 - t;
      const float c = v.C[doff(l - k, L) + k];
      float sc;
      if constexpr (CONTRA) sc = c2_twoloop_outer(T, s, o, k, l, a, b);
      else sc = t_twoloop(T, s, i, j, k, l, a, b);
      sum = lse(sum, __fadd_rn(c, sc), lut);
    }
    const float mb = (d >= 2) ? Mm2[i + 1] : NEG;
    sum = lse(sum, __fadd_rn(mb, v2_mbclose<CONTRA>(T, s, L, i, j)), lut);
    if (sum > NEG) v.C[od + i] = sum;
  }
}

// =========================================================================================================
// inside, role Y (CONTRAfold): the k < j part of sums_rightmost_basepairs_{external,multibranch}[i][j]
// (src/mccaskill_algo.rs:468-486); the k == j term is added by role Z one step later.
// Needs: sums_close of diagonals < d.
// =========================================================================================================
template <class SV>
RNA_DEV void inside_Y_contra(const SV& v, const ContraView2& T, const float4* lut, int d, int lane, int nl) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  const DevContra* dev = T.g;
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    float r = NEG, rm = NEG;
    const uint32_t* row = v.mask + i * v.W2;
    for (int p = i + 1; p <= j - 1; p += 32) {
      uint32_t w = get32(row, p);
      const int n = j - p;
      if (n < 32) w &= (1u << n) - 1u;
      while (w) {
        const int t = __ffs(w) - 1;
        w &= w - 1;
        const int k = p + t;
        const float av = __fadd_rn(v.C[doff(k - i, L) + i], v2_acc<true>(T, s, L, i, k));
        const float nn = (float)(j - k);
        r = lse(r, __fadd_rn(__fadd_rn(av, dev->ext_bp), __fmul_rn(dev->ext_unpair, nn)), lut);
        rm = lse(rm, __fadd_rn(__fadd_rn(av, dev->mb_bp), __fmul_rn(dev->mb_unpair, nn)), lut);
      }
    }
    v.R[od + i] = r;
    v.X[od + i] = rm;
  }
}

// =========================================================================================================
// inside, role Z: finish R (/Rm) of diagonal d with the k == j term, then the three dense chains
// sums_external, sums_multibranch, sums_1ormore_basepairs (src/mccaskill_algo.rs:344-374, 468-512).
// Needs: sums_close(d) (role X, previous step), partial R/Rm(d) (role Y, previous step), R/Rm/E/M1 of < d.
// =========================================================================================================
template <bool CONTRA, class SV>
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
    float sE, sM1, sM = NEG;
    if constexpr (CONTRA) {
      sE = __fmul_rn(dev->ext_unpair, (float)(d + 1));
      sM1 = Rmij;
    } else {
      sE = 0.f;
      sM1 = __fadd_rn(Rij, dev->coeff_num_branches);
    }
    sE = lse(sE, __fadd_rn(Rij, 0.f), lut);   // k = i: E[i][i-1] = 0 (lower triangle / literal 0)
    for (int m = 1; m < d; m++) {
      const float r = v.R[doff(d - m, L) + i + m];
      const float e = v.E[doff(m - 1, L) + i];
      const float m1 = v.M1[doff(m - 1, L) + i];
      sE = lse(sE, __fadd_rn(r, e), lut);
      if constexpr (CONTRA) {
        const float rm = v.X[doff(d - m, L) + i + m];
        sM1 = lse(sM1, __fadd_rn(rm, __fmul_rn(dev->mb_unpair, (float)m)), lut);
        sM = lse(sM, __fadd_rn(m1, rm), lut);
      } else {
        const float xx = __fadd_rn(r, dev->coeff_num_branches);
        sM1 = lse(sM1, xx, lut);
        sM = lse(sM, __fadd_rn(m1, xx), lut);
      }
    }
    v.E[od + i] = sE;
    Mcur[i] = sM;
    sM1 = lse(sM1, sM, lut);
    v.M1[od + i] = sM1;
  }
}

// =========================================================================================================
// outside, role Y: probs_multibranch / probs_multibranch2 of every cell of diagonal d
// (src/mccaskill_algo.rs:540-557, 641-661).  Needs: log P of diagonals > d.
// =========================================================================================================
template <bool CONTRA, class SV>
RNA_DEV void outside_Y(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut, int d,
                       int lane, int nl) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int ncell = L - d, od = doff(d, L);
  for (int i = lane; i < ncell; i += nl) {
    const int j = i + d;
    float pm = NEG, pm2 = NEG;
    const uint32_t* row = v.mask + i * v.W2;
    for (int p = j + 1; p < L; p += 32) {
      uint32_t w = get32(row, p);
      while (w) {
        const int t = __ffs(w) - 1;
        w &= w - 1;
        const int k = p + t, m = k - j;
        const int q = doff(k - i, L) + i;
        const float c = v.C[q], pv = v.E[q];
        const float x = __fsub_rn(__fadd_rn(pv, v2_mbclose<CONTRA>(T, s, L, i, k)), c);
        const float m1 = (m >= 2) ? v.M1[doff(m - 2, L) + j + 1] : NEG;
        pm = lse(pm, __fadd_rn(x, m1), lut);
        if constexpr (CONTRA) pm2 = lse(pm2, __fadd_rn(x, __fmul_rn(T.g->mb_unpair, (float)(m - 1))), lut);
        else pm2 = lse(pm2, x, lut);
      }
    }
    v.R[od + i] = pm;
    v.X[od + i] = pm2;
  }
}

// =========================================================================================================
// outside, role X: log P(i,j) of the closable cells of diagonal d (src/mccaskill_algo.rs:558-604, 662-719).
// Needs: log P, probs_multibranch, probs_multibranch2 of diagonals > d.
// =========================================================================================================
template <bool CONTRA, class SV>
RNA_DEV void outside_X(const SV& v, const typename Model2<CONTRA>::View& T, const float4* lut,
                       const ModelParams& P, float Z, int d, int lane, int nl) {
  const int L = v.L;
  const uint8_t* s = v.s;
  const float NEG = RNA_NEG_INF;
  const int cnt = v.pcnt[d], od = doff(d, L);
  const typename Model2<CONTRA>::Dev* dev = T.g;
  for (int r = lane; r < cnt; r += nl) {
    const int i = v.plist[od + r], j = i + d;
    const float Cij = v.C[od + i];
    if (!(Cij > NEG)) continue;
    const float Aij = __fadd_rn(Cij, v2_acc<CONTRA>(T, s, L, i, j));
    const float El = (i < 1) ? 0.f : v.E0[i - 1];
    const float Er = (j > L - 2) ? 0.f : v.EL[j + 1];
    float sm;
    if constexpr (CONTRA) sm = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(El, Er), Aij), dev->ext_bp), Z);
    else sm = __fsub_rn(__fadd_rn(__fadd_rn(El, Aij), Er), Z);
    // enclosing two-loops: k descending from i-1, l ascending from j+1
    C2Inner in;
    if constexpr (CONTRA) { in.js = c2_js(T, s, j, i); in.bp = T.sm->bp[s[i] * 4 + s[j]]; in.kl = s[i] * 4 + s[j]; }
    const int bcap = L - 2 - j;
    const int amax = (bcap >= 0) ? min(P.MAX2, i - 1) : -1;
    int a = -1, k = i;
    uint32_t w = 0;
    for (;;) {
      while (w == 0 && a < amax) {
        a++;
        k = i - 1 - a;
        const int n = min(P.MAX2 - a, bcap) + 1;   // 1..31 positions j+1 .. j+n
        w = get32(v.mask + k * v.W2, j + 1) & (0xffffffffu >> (32 - n));
      }
      if (w == 0) break;
      const int t = __ffs(w) - 1;
      w &= w - 1;
      const int l = j + 1 + t, b = t;
      const int q = doff(l - k, L) + k;
      const float c = v.C[q], pv = v.E[q];
      float tl;
      if constexpr (CONTRA) tl = c2_twoloop_inner(T, s, in, k, l, a, b);
      else tl = t_twoloop(T, s, k, l, i, j, a, b);
      sm = lse(sm, __fadd_rn(__fsub_rn(__fadd_rn(pv, Cij), c), tl), lut);
    }
    // enclosing multiloops: k ascending 0..i-1
    float sa;
    if constexpr (CONTRA) sa = __fadd_rn(Aij, dev->mb_bp); else sa = __fadd_rn(Aij, dev->coeff_num_branches);
    for (int kk = 0; kk < i; kk++) {
      const int m = i - 1 - kk;
      const int q = doff(j - kk, L) + kk;
      const float x = (m >= 1) ? v.M1[doff(m - 1, L) + kk + 1] : NEG;
      const float p2 = v.X[q], y = v.R[q];
      sm = lse(sm, __fadd_rn(__fadd_rn(sa, p2), x), lut);
      if constexpr (CONTRA) sm = lse(sm, __fadd_rn(__fadd_rn(sa, y), __fmul_rn(dev->mb_unpair, (float)m)), lut);
      else sm = lse(sm, __fadd_rn(sa, y), lut);
      sm = lse(sm, __fadd_rn(__fadd_rn(sa, x), y), lut);
    }
    if (sm > NEG) v.E[od + i] = sm;
  }
}

}  // namespace rna
