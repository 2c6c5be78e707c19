// fast_kernel.cuh — FAST numeric mode of the McCaskill inside / outside passes (north_star (2): "warp-shuffle
// log-sum-exp reductions over split points and interior loops").
//
// The reference folds every cell's terms one after the other with a piecewise-cubic, NON-ASSOCIATIVE logsumexp
// (src/utils.rs:579-627); the reference-exact kernels (fold_kernel2.cuh) must keep that order and are bound by the
// latency of the sequential chain.  This mode computes the same recurrences (src/mccaskill_algo.rs:282-723) in EXACT
// log-space arithmetic, which is associative up to rounding, so
//   * a cell is a WARP task: the 32 lanes stride over the split points k (and over the <= 496 interior-loop candidates
//     (a,b)), each lane keeps a running (max, sum of exp) pair, and the lanes meet in a 5-step shuffle reduction;
//   * every operand matrix is stored in the orientation in which its reduction walks it (R, Rm, probs_multibranch(2)
//     transposed; sums_1ormore in both orientations), so a warp's loads are 128-byte coalesced lines;
//   * the sums that the reference re-evaluates per cell but that obey a one-term recurrence are carried along:
//       R[i][j]   = lse(R[i][j-1] + unpair, A(i,j) + basepair)                 (src/mccaskill_algo.rs:344-351, 468-486)
//       S[i][j]   = lse(Rm[i][j], unpair + S[i+1][j])   (first chain of :364-374 / :499-512; Turner: Q[i][j] + COEFF)
//       PM2[i][j] = lse(unpair + PM2[i][j+1], x(i,j+1))                        (:540-557, 641-661)
//       T2[i][j]  = lse(PM[i-1][j], unpair + T2[i-1][j])                       (middle term of :594-601, 701-714)
//     which leaves two O(span) reductions per cell in each pass.
// exp / log: real = float -> ex2.approx / lg2.approx through __expf / __logf (MUFU pipe); real = double -> libdevice.
// One barrier per anti-diagonal: __syncthreads (one CTA per sequence, batches) or the grid barrier of the cooperative
// launch (one long sequence on the whole GPU).  Results differ from the reference's by its own approximation error
// (<= 7.6e-6 per logsumexp); tolerances are stated in DESIGN.md §2 and asserted in tests/test_fast_mode.py.
#pragma once
#include <math_constants.h>

#include "fold_kernel.cuh"
#include "fold_phases.cuh"

namespace rna {

struct FastArgs {
  const uint8_t* bases;
  const uint32_t* offsets;
  const uint32_t* order;
  uint32_t n_launch;
  const void* tables;          // DevTurner* / DevContra*
  int allows_short;
  float* out_logz;
  float* out_bpp;              // packed, RNA_BPP_ABSENT for absent keys
  const uint64_t* bpp_offsets;
  void* workspace;             // per CTA (batch) or one (cooperative): FAST_NMAT L x L matrices of `real` + vectors + codes
  unsigned long long ws_stride;   // bytes per slot
  int* work_counter;           // batch: dynamic work queue; cooperative: the grid barrier's counter
};

#define RNA_FAST_NMAT 7
__host__ __device__ inline size_t fast_slot_bytes(int L, size_t real_size) {
  const size_t LL = (size_t)L * (size_t)L;
  // matrices | 6 vectors of L | bases, RR, LL codes (padded) | candidate table
  return (RNA_FAST_NMAT * LL + 6 * (size_t)L + 8) * real_size + 4 * ((size_t)L + 16) + 1024 + 256;
}

template <class real> struct FastMath;
template <> struct FastMath<float> {
  static __device__ __forceinline__ float ex(float x) { return __expf(x); }     // ex2.approx(x * log2 e)
  static __device__ __forceinline__ float lg(float x) { return __logf(x); }     // lg2.approx(x) * ln 2
  static __device__ __forceinline__ float ninf() { return __int_as_float(0xff800000); }
};
template <> struct FastMath<double> {
  static __device__ __forceinline__ double ex(double x) { return exp(x); }
  static __device__ __forceinline__ double lg(double x) { return log(x); }
  static __device__ __forceinline__ double ninf() { return -CUDART_INF; }
};

// running log-sum-exp: value = m + log(s)
template <class real>
struct LseAcc {
  real m, s;
  __device__ __forceinline__ LseAcc() : m(FastMath<real>::ninf()), s(0) {}
  __device__ __forceinline__ void add(real x) {
    const real mn = (m > x) ? m : x;
    if (mn > FastMath<real>::ninf()) {
      s = s * FastMath<real>::ex(m - mn) + FastMath<real>::ex(x - mn);
      m = mn;
    }
  }
  __device__ __forceinline__ void merge(real m2, real s2) {
    const real mn = (m > m2) ? m : m2;
    if (mn > FastMath<real>::ninf()) {
      s = s * FastMath<real>::ex(m - mn) + s2 * FastMath<real>::ex(m2 - mn);
      m = mn;
    }
  }
  __device__ __forceinline__ real value() const { return (s > 0) ? m + FastMath<real>::lg(s) : FastMath<real>::ninf(); }
};
template <class real>
__device__ __forceinline__ real warp_lse(LseAcc<real> a) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const real m2 = __shfl_xor_sync(0xffffffffu, a.m, off);
    const real s2 = __shfl_xor_sync(0xffffffffu, a.s, off);
    a.merge(m2, s2);
  }
  return a.value();
}
// the lanes' partial sums merged (every lane holds the result), not yet a logarithm
template <class real>
__device__ __forceinline__ LseAcc<real> warp_merge(LseAcc<real> a) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const real m2 = __shfl_xor_sync(0xffffffffu, a.m, off);
    const real s2 = __shfl_xor_sync(0xffffffffu, a.s, off);
    a.merge(m2, s2);
  }
  return a;
}
template <class real>
__device__ __forceinline__ real lse2(real a, real b) {
  LseAcc<real> x;
  x.add(a);
  x.add(b);
  return x.value();
}

struct FastSeq {   // what the loop scorers of fold_phases.cuh read
  int L;
  const uint8_t* s;
  uint8_t* RR;
  uint8_t* LL;
};

#define RNA_FAST_NT_BATCH 256
#define RNA_FAST_NT_COOP 512   // one CTA per SM in the cooperative launch: half as many arrivals at the grid barrier
template <bool CONTRA, bool COOP, class real>
__global__ void __launch_bounds__(COOP ? RNA_FAST_NT_COOP : RNA_FAST_NT_BATCH, COOP ? 1 : 2) fast_fold_kernel(const FastArgs a) {
  typedef typename Model2<CONTRA>::Dev Dev;
  typedef typename Model2<CONTRA>::View View;
  typedef FastMath<real> FM;
  const Dev* dev = reinterpret_cast<const Dev*>(a.tables);
  View T;
  T.g = dev;
  T.sm = dev_small<CONTRA>(dev);   // (tables stay in global memory: read-only, L1-resident)
  ModelParams P;
  P.MINSPAN = dev->min_span;
  if constexpr (CONTRA) P.MAX2 = dev->max_loop_len; else P.MAX2 = dev->max_2loop_len;
  P.allows_short = a.allows_short;
  const real NEG = FM::ninf();
  const int tid = threadIdx.x, lane = tid & 31;
  // warps of the team that works on one sequence: the CTA (batch) or the whole grid (cooperative)
  const int gtid = COOP ? (int)(blockIdx.x * blockDim.x + tid) : tid;
  const int gnt = COOP ? (int)(gridDim.x * blockDim.x) : (int)blockDim.x;
  unsigned bar_target = 0;
  unsigned* const bar_ctr = reinterpret_cast<unsigned*>(a.work_counter);
  auto team_sync = [&]() { if constexpr (COOP) coop_barrier(bar_ctr, bar_target); else __syncthreads(); };
  // operands written by other SMs in earlier steps: read through L2 in the cooperative launch
  auto ld = [](const real* p) -> real { if constexpr (COOP) return __ldcg(p); else return *p; };
  __shared__ int s_work;
  __shared__ uint8_t cand[2 * 496];             // (a, b) of the interior-loop enumeration, a + b <= MAX2 <= 30
  // A diagonal with fewer cells than the team has warps (the LONG diagonals, whose cells carry the longest reductions)
  // gives each cell S warps of ONE CTA: every warp reduces its share of the split points / candidates, the partial
  // (max, sum) pairs meet in shared memory and the first warp of the cell finishes it.  Cells go round-robin over the
  // CTAs of the team (cell c: CTA c % G, slot c / G), so all CTAs see the same number of slots and rounds.
  __shared__ real s_red[(COOP ? RNA_FAST_NT_COOP : RNA_FAST_NT_BATCH) / 32][4][2];
  constexpr int UR = COOP ? 8 : 4;   // split points per lane in flight (the cooperative launch has 128 registers per thread)
  const int G = COOP ? (int)gridDim.x : 1, cta = COOP ? (int)blockIdx.x : 0;
  const int WPC = (int)(blockDim.x >> 5), wic = tid >> 5;
  int ncand = 0;
  for (int aa = 0; aa <= P.MAX2; aa++) ncand += P.MAX2 - aa + 1;
  for (int x = tid; x < ncand; x += blockDim.x) {
    int aa = 0, rest = x;
    while (rest >= P.MAX2 - aa + 1) { rest -= P.MAX2 - aa + 1; aa++; }
    cand[2 * x] = (uint8_t)aa; cand[2 * x + 1] = (uint8_t)rest;
  }
  __syncthreads();

  float mu = 0.f, eu = 0.f, ebp = 0.f, mbp = 0.f, cnb = 0.f;
  if constexpr (CONTRA) { mu = dev->mb_unpair; eu = dev->ext_unpair; ebp = dev->ext_bp; mbp = dev->mb_bp; }
  else cnb = dev->coeff_num_branches;

  for (uint32_t wloop = 0;; wloop++) {
    uint32_t w;
    if constexpr (COOP) {
      w = wloop;
    } else {
      __syncthreads();
      if (tid == 0) s_work = atomicAdd(a.work_counter, 1);
      __syncthreads();
      w = (uint32_t)s_work;
    }
    if (w >= a.n_launch) break;
    const uint32_t sidx = a.order ? a.order[w] : w;
    const uint32_t sbeg = a.offsets[sidx];
    const int L = (int)(a.offsets[sidx + 1] - sbeg);
    const size_t LL2 = (size_t)L * (size_t)L;

    // ---- carve the slot ---------------------------------------------------------------------------------------
    unsigned char* slot = reinterpret_cast<unsigned char*>(a.workspace) + (COOP ? 0 : (size_t)blockIdx.x * a.ws_stride);
    real* mC = reinterpret_cast<real*>(slot);     // sums_close, row-major [i][j]
    real* mE = mC + LL2;                          // sums_external [i][j]             -> outside: log P [i][j]
    real* mRT = mE + LL2;                         // R transposed [j][k] = R[k][j]    -> outside: probs_multibranch^T [j][k]
    real* mXT = mRT + LL2;                        // Rm transposed (CONTRAfold)       -> outside: probs_multibranch2^T [j][k]
    real* mM = mXT + LL2;                         // sums_multibranch [i][j]          -> outside: x(i,k) = P + mbclose - C
    real* mM1 = mM + LL2;                         // sums_1ormore_basepairs [i][j]
    real* mM1T = mM1 + LL2;                       // ... transposed [j][i]
    real* vS0 = mM1T + LL2;                       // rolling S (inside) / T2 (outside), by column j, two diagonals
    real* vS1 = vS0 + L;
    real* vE0 = vS1 + L;                          // sums_external[0][x]
    real* vEL = vE0 + L;                          // sums_external[x][L-1]
    uint8_t* sq = reinterpret_cast<uint8_t*>(vEL + 2 * L + 8) + 4;   // bases with zero pads on both sides
    uint8_t* RR = sq + L + 8;
    uint8_t* LLc = RR + L + 4;
    FastSeq v;
    v.L = L; v.s = sq; v.RR = RR; v.LL = LLc;

    // ---- setup --------------------------------------------------------------------------------------------------
    for (int x = gtid; x < L + 8; x += gnt) sq[x - 4] = (x >= 4 && x < L + 4) ? a.bases[sbeg + x - 4] : 0;
    for (size_t x = gtid; x < LL2; x += gnt) { mC[x] = NEG; mE[x] = 0; mRT[x] = NEG; mXT[x] = NEG; mM[x] = NEG; mM1[x] = NEG; mM1T[x] = NEG; }
    for (int x = gtid; x < 2 * L; x += gnt) vS0[x] = NEG;
    team_sync();
    for (int p = gtid; p < L; p += gnt) { RR[p] = (uint8_t)(sq[p] * 4 + sq[p + 1]); LLc[p] = (uint8_t)(sq[p] * 4 + sq[p - 1]); }
    team_sync();

    const int d_in0 = CONTRA ? 0 : (P.MINSPAN - 1);
    const int d_out0 = CONTRA ? (a.allows_short ? 1 : P.MINSPAN - 1) : (P.MINSPAN - 1);
    auto closable = [&](int i, int j) -> bool {
      if (!canonical_pair(sq[i], sq[j])) return false;
      if (CONTRA && P.allows_short) return true;
      return j - i + 1 >= P.MINSPAN;
    };

    // ================================ inside (src/mccaskill_algo.rs:282-516) ==================================
    for (int d = d_in0; d < L; d++) {
      real* Sprev = (d & 1) ? vS0 : vS1;   // S of diagonal d-1 (by column j)
      real* Scur = (d & 1) ? vS1 : vS0;
      const int ncell = L - d, cpc = (ncell + G - 1) / G;
      int S = 1;
      while (S * 2 * cpc <= WPC) S *= 2;                    // warps per cell (1 while the diagonal has cells for every warp)
      const int spr = WPC / S, rounds = (cpc + spr - 1) / spr;   // slots per round, rounds (identical in every CTA)
      const int part = wic % S;
      for (int rd = 0; rd < rounds; rd++) {
        const int slot = rd * spr + wic / S;
        const int i = slot * G + cta, j = i + d;
        const bool active = slot < cpc && i < ncell;
        const bool clos = active && closable(i, j);
        // ---- partial sums of this warp's share: interior-loop candidates, the two O(span) reductions ----
        LseAcc<real> acc, e, m;
        if (clos) {
          const typename LoopOf<CONTRA, true>::type lp = make_loop<CONTRA, true>(v, T, i, j);
          for (int c0 = lane + 128 * part; c0 < ncand; c0 += 128 * S) {   // four candidates per lane in flight
            real ck[4];
            int ca[4], cb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int c = c0 + 32 * u;
              ck[u] = NEG; ca[u] = 0; cb[u] = 0;
              if (c < ncand) {
                ca[u] = cand[2 * c]; cb[u] = cand[2 * c + 1];
                const int k = i + 1 + ca[u], l = j - 1 - cb[u];
                if (k < l) ck[u] = ld(&mC[(size_t)k * L + l]);
              }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              if (ck[u] > NEG) {
                const int k = i + 1 + ca[u], l = j - 1 - cb[u];
                Term1 t1;
                t1.c = 0.f; t1.pv = 0.f; t1.q = 0; t1.a = ca[u]; t1.b = cb[u]; t1.code = RR[l] * 16 + LLc[k];
                acc.add(ck[u] + (real)lp.score(lp.stage2(t1)));
              }
            }
          }
        }
        if (active) {
          const real* rowE = mE + (size_t)i * L;
          const real* rowM1 = mM1 + (size_t)i * L;
          const real* colR = mRT + (size_t)j * L;
          const real* colX = mXT + (size_t)j * L;
          for (int k0 = i + 1 + lane + 32 * UR * part; k0 <= j - 1; k0 += 32 * UR * S) {   // UR split points per lane in flight
            real r[UR], ee[UR], mm[UR], xx[UR];
#pragma unroll
            for (int u = 0; u < UR; u++) {
              const int k = k0 + 32 * u;
              r[u] = NEG; ee[u] = NEG; mm[u] = NEG; xx[u] = NEG;
              if (k <= j - 1) {
                r[u] = ld(&colR[k]); ee[u] = ld(&rowE[k - 1]); mm[u] = ld(&rowM1[k - 1]);
                if constexpr (CONTRA) xx[u] = ld(&colX[k]);
              }
            }
#pragma unroll
            for (int u = 0; u < UR; u++) {
              e.add(r[u] + ee[u]);
              if constexpr (CONTRA) m.add(mm[u] + xx[u]);
              else m.add(mm[u] + (r[u] + (real)cnb));
            }
          }
        }
        acc = warp_merge(acc); e = warp_merge(e); m = warp_merge(m);
        if (S > 1) {   // (uniform in the CTA)
          if (lane == 0) {
            s_red[wic][0][0] = acc.m; s_red[wic][0][1] = acc.s; s_red[wic][1][0] = e.m; s_red[wic][1][1] = e.s;
            s_red[wic][2][0] = m.m; s_red[wic][2][1] = m.s;
          }
          __syncthreads();
          if (part == 0) {
            for (int q = 1; q < S; q++) {
              acc.merge(s_red[wic + q][0][0], s_red[wic + q][0][1]);
              e.merge(s_red[wic + q][1][0], s_red[wic + q][1][1]);
              m.merge(s_red[wic + q][2][0], s_red[wic + q][2][1]);
            }
          }
          __syncthreads();
        }
        if (!active || part != 0) continue;
        // ---- the cell's first warp finishes it (every lane computes the same scalars, lane 0 stores) ----
        // (1) sums_close: hairpin (+) interior loops (+) multiloop closing
        real Cij = NEG;
        if (clos) {
          if constexpr (CONTRA) { if (d - 1 <= P.MAX2) acc.add((real)c2_hairpin(T, sq, i, j)); }
          else acc.add((real)t_hairpin(T, sq, i, j));
          if (d >= 2) acc.add(ld(&mM[(size_t)(i + 1) * L + (j - 1)]) + (real)v2_mbclose<CONTRA>(T, sq, L, i, j));
          Cij = acc.value();
        }
        // (2) rightmost-pair sums by their one-term recurrences
        const real Aij = (Cij > NEG) ? Cij + (real)v2_acc<CONTRA>(T, sq, L, i, j) : NEG;
        const real Rprev = (d >= 1) ? ld(&mRT[(size_t)(j - 1) * L + i]) : NEG;
        real Rij, Rmij = NEG;
        if constexpr (CONTRA) {
          const real Xprev = (d >= 1) ? ld(&mXT[(size_t)(j - 1) * L + i]) : NEG;
          Rij = lse2<real>(Rprev + (real)eu, Aij + (real)ebp);
          Rmij = lse2<real>(Xprev + (real)mu, Aij + (real)mbp);
        } else {
          Rij = lse2<real>(Rprev, Aij);
        }
        // (3) sums_external, (4) sums_multibranch: the split-point reductions plus the k = i terms
        e.add(CONTRA ? (real)eu * (real)(d + 1) : (real)0);
        e.add(Rij);                                     // k = i: E[i][i-1] = 0
        const real Eij = e.value(), Mij = m.value();
        // first chain of the 1-or-more sum by its recurrence down the column
        real Sij;
        if constexpr (CONTRA) Sij = lse2<real>(Rmij, (real)mu + ld(&Sprev[j]));
        else Sij = lse2<real>(Rij, ld(&Sprev[j]));
        const real M1ij = lse2<real>(CONTRA ? Sij : Sij + (real)cnb, Mij);
        if (lane == 0) {
          if (clos) mC[(size_t)i * L + j] = Cij;
          mRT[(size_t)j * L + i] = Rij;
          if constexpr (CONTRA) mXT[(size_t)j * L + i] = Rmij;
          mE[(size_t)i * L + j] = Eij;
          mM[(size_t)i * L + j] = Mij;
          mM1[(size_t)i * L + j] = M1ij;
          mM1T[(size_t)j * L + i] = M1ij;
          Scur[j] = Sij;
        }
      }
      team_sync();
    }

    // ================================ outside (src/mccaskill_algo.rs:518-723) =================================
    for (int x = gtid; x < L; x += gnt) { vE0[x] = ld(&mE[x]); vEL[x] = ld(&mE[(size_t)x * L + (L - 1)]); }
    team_sync();
    const real Z = ld(&vE0[L - 1]);
    for (size_t x = gtid; x < LL2; x += gnt) { mE[x] = NEG; mRT[x] = NEG; mXT[x] = NEG; mM[x] = NEG; }
    for (int x = gtid; x < 2 * L; x += gnt) vS0[x] = NEG;
    if (gtid == 0 && a.out_logz) a.out_logz[sidx] = (float)Z;
    team_sync();
    real* mP = mE;      // log P [i][j]
    real* mPMT = mRT;   // probs_multibranch^T [j][k]
    real* mPM2T = mXT;  // probs_multibranch2^T [j][k]
    real* mXQ = mM;     // x(i,k) = log P(i,k) + mbclose(i,k) - sums_close(i,k), row-major
    for (int d = L - 1; d >= d_out0; d--) {
      real* Tprev = (d & 1) ? vS0 : vS1;   // T2 of diagonal d+1 (by column j)
      real* Tcur = (d & 1) ? vS1 : vS0;
      const int ncell = L - d, cpc = (ncell + G - 1) / G;
      int S = 1;
      while (S * 2 * cpc <= WPC) S *= 2;
      const int spr = WPC / S, rounds = (cpc + spr - 1) / spr;
      const int part = wic % S;
      for (int rd = 0; rd < rounds; rd++) {
        const int slot = rd * spr + wic / S;
        const int i = slot * G + cta, j = i + d;
        const bool active = slot < cpc && i < ncell;
        const real Cij = active ? ld(&mC[(size_t)i * L + j]) : NEG;
        const bool clos = Cij > NEG;
        // ---- partial sums of this warp's share ----
        LseAcc<real> pmacc, acc, t1acc, t3acc;
        if (active) {   // probs_multibranch: reduction over k > j
          const real* rowXQ = mXQ + (size_t)i * L;
          const real* rowM1 = mM1 + (size_t)(j + 1) * L;
          for (int k0 = j + 2 + lane + 32 * UR * part; k0 < L; k0 += 32 * UR * S) {
            real xq4[UR], m14[UR];
#pragma unroll
            for (int u = 0; u < UR; u++) {
              const int k = k0 + 32 * u;
              xq4[u] = NEG; m14[u] = NEG;
              if (k < L) { xq4[u] = ld(&rowXQ[k]); m14[u] = ld(&rowM1[k - 1]); }
            }
#pragma unroll
            for (int u = 0; u < UR; u++) pmacc.add(xq4[u] + m14[u]);
          }
        }
        if (clos) {
          const typename LoopOf<CONTRA, false>::type lp = make_loop<CONTRA, false>(v, T, i, j);
          for (int c0 = lane + 128 * part; c0 < ncand; c0 += 128 * S) {
            real ck[4], pk[4];
            int ca[4], cb[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
              const int c = c0 + 32 * u;
              ck[u] = NEG; pk[u] = NEG; ca[u] = 0; cb[u] = 0;
              if (c < ncand) {
                ca[u] = cand[2 * c]; cb[u] = cand[2 * c + 1];
                const int k = i - 1 - ca[u], l = j + 1 + cb[u];
                if (k >= 0 && l < L) { ck[u] = ld(&mC[(size_t)k * L + l]); pk[u] = ld(&mP[(size_t)k * L + l]); }
              }
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
              if (ck[u] > NEG) {
                const int k = i - 1 - ca[u], l = j + 1 + cb[u];
                Term1 t1;
                t1.c = 0.f; t1.pv = 0.f; t1.q = 0; t1.a = ca[u]; t1.b = cb[u]; t1.code = RR[k] * 16 + LLc[l];
                acc.add(pk[u] + (Cij - ck[u]) + (real)lp.score(lp.stage2(t1)));
              }
            }
          }
          // enclosing multiloops: T1 = lse_k PM2[k][j] + M1[k+1][i-1], T3 = lse_k PM[k][j] + M1[k+1][i-1], k <= i-2
          if (i >= 2) {
            const real* colPM2 = mPM2T + (size_t)j * L;
            const real* colPM = mPMT + (size_t)j * L;
            const real* colM1 = mM1T + (size_t)(i - 1) * L;
            for (int k0 = lane + 32 * UR * part; k0 <= i - 2; k0 += 32 * UR * S) {
              real x14[UR], p24[UR], p4[UR];
#pragma unroll
              for (int u = 0; u < UR; u++) {
                const int k = k0 + 32 * u;
                x14[u] = NEG; p24[u] = NEG; p4[u] = NEG;
                if (k <= i - 2) { x14[u] = ld(&colM1[k + 1]); p24[u] = ld(&colPM2[k]); p4[u] = ld(&colPM[k]); }
              }
#pragma unroll
              for (int u = 0; u < UR; u++) { t1acc.add(p24[u] + x14[u]); t3acc.add(p4[u] + x14[u]); }
            }
          }
        }
        pmacc = warp_merge(pmacc); acc = warp_merge(acc); t1acc = warp_merge(t1acc); t3acc = warp_merge(t3acc);
        if (S > 1) {
          if (lane == 0) {
            s_red[wic][0][0] = pmacc.m; s_red[wic][0][1] = pmacc.s; s_red[wic][1][0] = acc.m; s_red[wic][1][1] = acc.s;
            s_red[wic][2][0] = t1acc.m; s_red[wic][2][1] = t1acc.s; s_red[wic][3][0] = t3acc.m; s_red[wic][3][1] = t3acc.s;
          }
          __syncthreads();
          if (part == 0) {
            for (int q = 1; q < S; q++) {
              pmacc.merge(s_red[wic + q][0][0], s_red[wic + q][0][1]);
              acc.merge(s_red[wic + q][1][0], s_red[wic + q][1][1]);
              t1acc.merge(s_red[wic + q][2][0], s_red[wic + q][2][1]);
              t3acc.merge(s_red[wic + q][3][0], s_red[wic + q][3][1]);
            }
          }
          __syncthreads();
        }
        if (!active || part != 0) continue;
        // ---- the cell's first warp finishes it ----
        // probs_multibranch2 by its recurrence, probs_multibranch from the reduction
        real pm2 = NEG;
        if (j + 1 < L) pm2 = lse2<real>(ld(&mPM2T[(size_t)(j + 1) * L + i]) + (real)mu, ld(&mXQ[(size_t)i * L + (j + 1)]));
        const real pm = pmacc.value();
        // T2[i][j] = lse over k < i of probs_multibranch[k][j] + unpair * (i - k - 1)
        real t2 = NEG;
        if (i >= 1) t2 = lse2<real>(ld(&mPMT[(size_t)j * L + (i - 1)]), (real)mu + ld(&Tprev[j]));
        // log P(i,j)
        real Pij = NEG, xq = NEG;
        if (clos) {
          const real Aij = Cij + (real)v2_acc<CONTRA>(T, sq, L, i, j);
          const real El = (i < 1) ? (real)0 : ld(&vE0[i - 1]), Er = (j > L - 2) ? (real)0 : ld(&vEL[j + 1]);
          acc.add(CONTRA ? El + Er + Aij + (real)ebp - Z : El + Aij + Er - Z);
          const real sa = Aij + (CONTRA ? (real)mbp : (real)cnb);
          acc.add(sa + t1acc.value()); acc.add(sa + t2); acc.add(sa + t3acc.value());
          Pij = acc.value();
          xq = Pij + (real)v2_mbclose<CONTRA>(T, sq, L, i, j) - Cij;
        }
        if (lane == 0) {
          mPMT[(size_t)j * L + i] = pm;
          mPM2T[(size_t)j * L + i] = pm2;
          Tcur[j] = t2;
          mP[(size_t)i * L + j] = Pij;
          mXQ[(size_t)i * L + j] = xq;
        }
      }
      team_sync();
    }
    // ================================ BPP = exp(log P), packed row-major ======================================
    if (a.out_bpp) {
      float* ob = a.out_bpp + a.bpp_offsets[sidx];
      for (int i = COOP ? (int)blockIdx.x : 0; i < L - 1; i += COOP ? (int)gridDim.x : 1) {
        const size_t rowoff = (size_t)i * (size_t)(2 * L - i - 1) / 2;
        for (int x = tid; x < L - 1 - i; x += blockDim.x) {
          const real p = ld(&mP[(size_t)i * L + (i + 1 + x)]);
          real pr = -1;
          if (p > NEG) { if constexpr (sizeof(real) == 8) pr = exp(p); else pr = expf(p); }
          ob[rowoff + x] = (float)pr;
        }
      }
    }
    team_sync();
  }
}

}  // namespace rna
