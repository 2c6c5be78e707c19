// fold_kernel.cuh — McCaskill inside/outside + BPP + centroid estimator as an anti-diagonal wavefront.
//
// One thread per DP cell of the current diagonal ("span"); every cell's fold over split points /
// interior loops is evaluated sequentially IN THE REFERENCE'S ORDER with the reference's polynomial
// logsumexp, so all values are bit-identical to the reference algorithm.
// All matrices are stored DIAGONAL-MAJOR (index(i,j) = off(j-i) + i): the cells of one diagonal are
// contiguous, so for every operand of every recurrence the 32 lanes of a warp (cells i..i+31 of the same
// diagonal) touch 32 consecutive words — conflict-free in shared memory, fully coalesced in HBM/L2.
// Loop-type branches (stack / bulge / 1x1 / ... / generic interior) depend only on the offsets
// (a,b) = (k-i-1, j-l-1), which are identical for all lanes: the scoring code is warp-uniform.
//
// Three storage/synchronisation modes share this one body:
//   MODE_SMEM   one CTA per sequence, matrices resident in shared memory      (tRNA .. ~145 nt)
//   MODE_GLOBAL one CTA per sequence, matrices in an HBM/L2 workspace slot    (Rfam-length batches)
//   MODE_COOP   the whole grid works on one sequence, grid-wide barrier per diagonal (1k-4k nt)
//
// Reference recurrences: src/mccaskill_algo.rs:282-378 (Turner inside), :380-516 (CONTRAfold inside),
// :518-610 / :612-723 (outside + BPP), src/centroid_fold.rs:25-105 (centroid + traceback),
// scorers src/utils.rs:166-556.
#pragma once
#include <cooperative_groups.h>

#include "dev_tables.h"
#include "numerics.cuh"
#include "scorers.cuh"

namespace rna {
namespace cg = cooperative_groups;

enum { MODE_SMEM = 0, MODE_GLOBAL = 1, MODE_COOP = 2 };

struct FoldArgs {
  const uint8_t* bases;
  const uint32_t* offsets;
  const uint32_t* order;     // sequence indices handled by this launch (length n_launch)
  uint32_t n_launch;
  uint32_t n_seqs;           // whole batch (stride of the per-gamma outputs)
  uint32_t total_len;
  int allows_short;
  const void* tables;        // DevTurner* / DevContra*
  const float* gammas;
  uint32_t n_gammas;
  float* out_logz;
  float* out_bpp;
  const uint64_t* bpp_offsets;
  uint8_t* out_structs;
  float* out_ea;
  uint16_t* out_pairs;
  uint32_t* out_npairs;
  float* workspace;          // MODE_GLOBAL / MODE_COOP
  unsigned long long ws_stride;   // floats per workspace slot
  int* work_counter;         // MODE_SMEM / MODE_GLOBAL: dynamic work queue
  int Lcap;                  // capacity the shared-memory carve-up was sized for
  int nXw, nYw, nZw;         // fold_kernel2: warps per role (threads beyond them only help in the all-thread phases)
  unsigned char* stream_ws;  // fold_kernel2: per-CTA slots for the two-loop term streams (null: score on the fly)
  unsigned long long stream_stride;   // bytes per slot
  uint32_t tcap;             // terms per stream a slot can hold
  long long* dbg;            // optional per-step per-warp cycle counts of CTA 0's first sequence (RNA_FOLD_DBG)
  int no_ml_split;           // cooperative kernel: A/B switch, multiloop chain without the producer warp
};

template <int MODE>
struct Ctx {
  typedef typename std::conditional<MODE == MODE_SMEM, int, long long>::type ofs_t;
  __device__ __forceinline__ static void sync() {
    if (MODE == MODE_COOP) cg::this_grid().sync(); else __syncthreads();
  }
  __device__ __forceinline__ static int first() {
    return MODE == MODE_COOP ? (int)(blockIdx.x * blockDim.x + threadIdx.x) : (int)threadIdx.x;
  }
  __device__ __forceinline__ static int stride() {
    return MODE == MODE_COOP ? (int)(gridDim.x * blockDim.x) : (int)blockDim.x;
  }
  // In MODE_COOP other SMs produce the data: read through L2 (ld.global.cg), never a stale L1 line.
  __device__ __forceinline__ static float ld(const float* p) { return MODE == MODE_COOP ? __ldcg(p) : *p; }
  __device__ __forceinline__ static ofs_t off(int d, int L) {
    return (ofs_t)d * (ofs_t)L - (((ofs_t)d * (ofs_t)(d - 1)) >> 1);
  }
};

// Shared-memory footprint (bytes) of the fixed part and of the SMEM-mode matrices for capacity Lcap.
template <bool CONTRA>
__host__ __device__ inline size_t fold_smem_fixed_bytes(int Lcap) {
  size_t b = 128;                                            // LSE coefficient LUT
  b += (sizeof(typename ModelTraits<CONTRA>::Small) + 15) / 16 * 16;
  b += ((size_t)Lcap + 8 + 15) / 16 * 16;                    // sequence bytes (+ guard)
  return b;
}
template <bool CONTRA>
__host__ __device__ inline size_t fold_ws_floats(int L) {
  const size_t T = (size_t)L * ((size_t)L + 1) / 2;
  const size_t nm = CONTRA ? 6 : 5;
  // matrices | Mroll 3L | E0 L | EL L | traceback stack 2(L+2) ints
  return nm * T + 3 * (size_t)L + 2 * (size_t)L + 2 * ((size_t)L + 2) + 8;
}
template <bool CONTRA>
__host__ __device__ inline size_t fold_smem_bytes(int Lcap, bool smem_mats) {
  size_t b = fold_smem_fixed_bytes<CONTRA>(Lcap);
  if (smem_mats) b += fold_ws_floats<CONTRA>(Lcap) * 4;
  return b;
}

// ---------------------------------------------------------------------------------------------------
// centroid_fold: src/centroid_fold.rs:25-105.  W (max_expect_accuracies) is diagonal-major in `W`;
// getp(d, i) returns the base-pairing probability of (i, i+d) or -1 when the key is absent.
// ---------------------------------------------------------------------------------------------------
// Traceback of one threshold (src/centroid_fold.rs:65-102) by ONE warp (all 32 lanes call it); the k-scan is
// lane-parallel (first match wins).  W, tstack: this threshold's matrix and stack.
template <int MODE, class PF>
__device__ __forceinline__ void centroid_traceback(const FoldArgs& a, uint32_t sidx, uint32_t sbeg, int L, const float* W,
                                                   int* tstack, PF getp, uint32_t g) {
  typedef Ctx<MODE> X;
  const float gamma = a.gammas[g];
  const int tid = threadIdx.x & 31;
  uint8_t* ostr = a.out_structs ? a.out_structs + (size_t)g * a.total_len + sbeg : nullptr;
  {
    const int lane = tid;
    int sp = 0;
    uint32_t np = 0;
    uint16_t* opairs = a.out_pairs ? a.out_pairs + 2 * ((size_t)g * a.total_len + sbeg) : nullptr;
    if (lane == 0) { tstack[0] = 0; tstack[1] = L - 1; }
    sp = 1;
    __syncwarp();
    while (sp > 0) {
      sp--;
      const int i = tstack[2 * sp], j = tstack[2 * sp + 1];
      __syncwarp();
      if (j <= i) continue;
      const int d = j - i;
      const float wv = X::ld(&W[X::off(d, L) + i]);
      if (wv == 0.f) continue;
      const float wi1 = X::ld(&W[X::off(d - 1, L) + i + 1]);
      const float wj1 = X::ld(&W[X::off(d - 1, L) + i]);
      const float p = getp(d, i);
      const float inner = (d >= 2) ? X::ld(&W[X::off(d - 2, L) + i + 1]) : 0.f;
      if (wv == wi1) {
        if (lane == 0) { tstack[2 * sp] = i + 1; tstack[2 * sp + 1] = j; }
        sp++;
      } else if (wv == wj1) {
        if (lane == 0) { tstack[2 * sp] = i; tstack[2 * sp + 1] = j - 1; }
        sp++;
      } else if (p != -1.0f && wv == __fsub_rn(__fadd_rn(inner, __fmul_rn(gamma, p)), 1.0f)) {
        if (lane == 0) {
          tstack[2 * sp] = i + 1; tstack[2 * sp + 1] = j - 1;
          if (ostr) { ostr[i] = '('; ostr[j] = ')'; }
          if (opairs) { opairs[2 * np] = (uint16_t)i; opairs[2 * np + 1] = (uint16_t)j; }
        }
        sp++;
        np++;
      } else {
        for (int m0 = 1; m0 < d; m0 += 32) {
          const int m = m0 + lane;
          bool hit = false;
          if (m < d) {
            hit = (wv == __fadd_rn(X::ld(&W[X::off(m, L) + i]), X::ld(&W[X::off(d - m - 1, L) + i + m + 1])));
          }
          const unsigned bal = __ballot_sync(0xffffffffu, hit);
          if (bal) {
            const int k = i + m0 + (__ffs(bal) - 1);
            if (lane == 0) {
              tstack[2 * sp] = i; tstack[2 * sp + 1] = k;
              tstack[2 * sp + 2] = k + 1; tstack[2 * sp + 3] = j;
            }
            sp += 2;
            break;
          }
        }
      }
      __syncwarp();
    }
    if (lane == 0) {
      if (a.out_ea) a.out_ea[(size_t)g * a.n_seqs + sidx] = X::ld(&W[X::off(L - 1, L)]);
      if (a.out_npairs) a.out_npairs[(size_t)g * a.n_seqs + sidx] = np;
    }
  }
}

struct NoCentroidFill { __device__ __forceinline__ bool operator()(float) const { return false; } };
// fill(gamma): optional replacement of the max-plus fill (W is zeroed and a barrier has passed; it must end with a
// barrier); returns false to get the built-in thread-per-cell fill.
template <int MODE, class PF, class FILL = NoCentroidFill>
__device__ __forceinline__ void centroid_run(const FoldArgs& a, uint32_t sidx, uint32_t sbeg, int L, float* W,
                                             int* tstack, PF getp, FILL fill = FILL()) {
  typedef Ctx<MODE> X;
  typedef typename X::ofs_t ofs_t;
  const int tid = threadIdx.x;
  const int c0 = X::first(), cs = X::stride();
  const ofs_t TRI = (ofs_t)L * (ofs_t)(L + 1) / 2;
  for (uint32_t g = 0; g < a.n_gammas; g++) {
    const float gamma = a.gammas[g];
    for (ofs_t x = c0; x < TRI; x += cs) W[x] = 0.f;
    uint8_t* ostr = a.out_structs ? a.out_structs + (size_t)g * a.total_len + sbeg : nullptr;
    if (ostr) for (int x = c0; x < L; x += cs) ostr[x] = '.';
    X::sync();
    if (!fill(gamma))
    for (int d = 1; d < L; d++) {
      const int ncell = L - d;
      const ofs_t od = X::off(d, L);
      const ofs_t od1 = X::off(d - 1, L);
      for (int i = c0; i < ncell; i += cs) {
        float wv = X::ld(&W[od1 + i + 1]);
        float e = X::ld(&W[od1 + i]);
        if (e > wv) wv = e;
        const float p = getp(d, i);
        if (p != -1.0f) {
          const float inner = (d >= 2) ? X::ld(&W[X::off(d - 2, L) + i + 1]) : 0.f;
          e = __fsub_rn(__fadd_rn(inner, __fmul_rn(gamma, p)), 1.0f);
          if (e > wv) wv = e;
        }
#pragma unroll 8
        for (int m = 1; m < d; m++) {   // (loads are independent of the running maximum: unrolled, they overlap)
          e = __fadd_rn(X::ld(&W[X::off(m, L) + i]), X::ld(&W[X::off(d - m - 1, L) + i + m + 1]));
          if (e > wv) wv = e;
        }
        W[od + i] = wv;
      }
      X::sync();
    }
    // traceback by the first warp of the (first) CTA
    if ((MODE != MODE_COOP || blockIdx.x == 0) && tid < 32) centroid_traceback<MODE>(a, sidx, sbeg, L, W, tstack, getp, g);
    X::sync();
  }
}

// ---------------------------------------------------------------------------------------------------
// The kernel
// ---------------------------------------------------------------------------------------------------
template <bool CONTRA, int MODE>
__global__ void __launch_bounds__(MODE == MODE_SMEM ? 256 : 512) fold_kernel(const FoldArgs a) {
  typedef Ctx<MODE> X;
  typedef typename X::ofs_t ofs_t;
  typedef typename ModelTraits<CONTRA>::Dev Dev;
  typedef typename ModelTraits<CONTRA>::Small Small;
  typedef typename ModelTraits<CONTRA>::View View;

  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* lut = reinterpret_cast<float4*>(smem_raw);
  Small* small = reinterpret_cast<Small*>(smem_raw + 128);
  uint8_t* sseq = smem_raw + 128 + (sizeof(Small) + 15) / 16 * 16;
  float* smats = reinterpret_cast<float*>(smem_raw + fold_smem_fixed_bytes<CONTRA>(a.Lcap));
  __shared__ int s_work;

  const Dev* dev = reinterpret_cast<const Dev*>(a.tables);
  const int tid = threadIdx.x;
  load_lse_lut(lut);
  for (int x = tid; x < (int)(sizeof(Small) / 4); x += blockDim.x)
    reinterpret_cast<float*>(small)[x] = reinterpret_cast<const float*>(&dev->small)[x];
  View T;
  T.g = dev;
  T.sm = small;
  __syncthreads();

  const float NEG = RNA_NEG_INF;
  const int MINSPAN = dev->min_span;
  int MAX2;
  if constexpr (CONTRA) MAX2 = dev->max_loop_len; else MAX2 = dev->max_2loop_len;

  for (uint32_t wloop = 0;; wloop++) {
    uint32_t w;
    if (MODE == MODE_COOP) {
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
    const ofs_t TRI = (ofs_t)L * (ofs_t)(L + 1) / 2;

    float* base;
    if (MODE == MODE_SMEM) base = smats;
    else if (MODE == MODE_GLOBAL) base = a.workspace + (size_t)blockIdx.x * a.ws_stride;
    else base = a.workspace;
    float* mC = base;
    float* mR = mC + TRI;       // R -> PM -> W (centroid)
    float* mE = mR + TRI;       // E -> P (log) -> prob
    float* mM1 = mE + TRI;
    float* mX = mM1 + TRI;      // Turner: PM2 (outside only); CONTRAfold: Rm -> PM2
    float* mA = mX + TRI;       // CONTRAfold only: A
    float* Mroll = CONTRA ? mA + TRI : mX + TRI;   // sums_multibranch: only diagonals d, d-1, d-2 are live
    float* E0 = Mroll + 3 * (size_t)L;             // sums_external[0][x]
    float* EL = E0 + L;                            // sums_external[x][L-1]
    int* tstack = reinterpret_cast<int*>(EL + L);

    uint8_t* s = sseq + 4;
    for (int x = tid; x < L; x += blockDim.x) s[x] = a.bases[sbeg + x];
    if (tid < 4) { sseq[tid] = 0; s[L + tid] = 0; }

    const int c0 = X::first(), cs = X::stride();
    // ---- init (FoldSums::new, src/mccaskill_algo.rs:213-226) ---------------------------------------
    for (ofs_t x = c0; x < TRI; x += cs) {
      mC[x] = NEG; mR[x] = NEG; mE[x] = 0.f; mM1[x] = NEG; mX[x] = NEG;
      if (CONTRA) mA[x] = NEG;
    }
    for (int x = c0; x < 3 * L; x += cs) Mroll[x] = NEG;
    X::sync();

    // ================================ inside =======================================================
    const int d_in0 = CONTRA ? 0 : (MINSPAN - 1);
    for (int d = d_in0; d < L; d++) {
      const int ncell = L - d;
      const ofs_t od = X::off(d, L);
      float* Mcur = Mroll + (size_t)(d % 3) * L;
      const float* Mm2 = Mroll + (size_t)((d + 1) % 3) * L;   // diagonal d-2
      for (int i = c0; i < ncell; i += cs) {
        const int j = i + d;
        const int si = s[i], sj = s[j];
        bool pairable = canonical_pair(si, sj);
        if (CONTRA) pairable = pairable && (a.allows_short || d + 1 >= MINSPAN);
        float sumC = NEG;
        // (1) sums_close
        if (__any_sync(__activemask(), pairable)) {
          if (pairable) {
            if constexpr (CONTRA) {
              if (d - 1 <= MAX2) sumC = lse(sumC, c_hairpin(T, s, i, j), lut);
            } else {
              sumC = lse(sumC, t_hairpin(T, s, i, j), lut);
            }
          }
          const int amax = min(MAX2, d - 3);
          for (int aa = 0; aa <= amax; aa++) {
            const int k = i + 1 + aa;
            const int bmax = min(MAX2 - aa, d - 3 - aa);
            for (int bb = 0; bb <= bmax; bb++) {
              const int l = j - 1 - bb;
              const float c = X::ld(&mC[X::off(d - 2 - aa - bb, L) + k]);
              const bool on = pairable && (c > NEG);
              if (__any_sync(__activemask(), on)) {
                const float y = __fadd_rn(c, m_twoloop<CONTRA>(T, s, i, j, k, l, aa, bb));
                sumC = lse_if(on, sumC, y, lut);
              }
            }
          }
          if (pairable) {
            const float mbc = m_mbclose<CONTRA>(T, s, L, i, j);
            const float mb = (d >= 2) ? X::ld(&Mm2[i + 1]) : NEG;
            sumC = lse(sumC, __fadd_rn(mb, mbc), lut);
          }
        }
        float accv = NEG;
        if (sumC > NEG) {
          mC[od + i] = sumC;
          accv = __fadd_rn(sumC, m_acc<CONTRA>(T, s, L, i, j));
          if (CONTRA) mA[od + i] = accv;
        }
        // (2) sums_rightmost_basepairs_external (/ _multibranch)
        float Rij, Rmij = NEG;
        if constexpr (!CONTRA) {
          // prefix property of the left-to-right fold: R[i][j] = R[i][j-1] (+) A(i,j)
          const float prev = (d >= 1) ? X::ld(&mR[X::off(d - 1, L) + i]) : NEG;
          Rij = lse(prev, accv, lut);
        } else {
          Rij = NEG;
          for (int m = 1; m <= d; m++) {
            const float av = (m == d) ? accv : X::ld(&mA[X::off(m, L) + i]);
            const float n = (float)(d - m);
            Rij = lse(Rij, __fadd_rn(__fadd_rn(av, dev->ext_bp), __fmul_rn(dev->ext_unpair, n)), lut);
            Rmij = lse(Rmij, __fadd_rn(__fadd_rn(av, dev->mb_bp), __fmul_rn(dev->mb_unpair, n)), lut);
          }
          mX[od + i] = Rmij;
        }
        mR[od + i] = Rij;
        // (3) sums_external, (4) sums_multibranch / sums_1ormore_basepairs — three independent chains
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
          const float r = X::ld(&mR[X::off(d - m, L) + i + m]);
          const float e = X::ld(&mE[X::off(m - 1, L) + i]);
          const float m1 = X::ld(&mM1[X::off(m - 1, L) + i]);
          sE = lse(sE, __fadd_rn(r, e), lut);
          if constexpr (CONTRA) {
            const float rm = X::ld(&mX[X::off(d - m, L) + i + m]);
            sM1 = lse(sM1, __fadd_rn(rm, __fmul_rn(dev->mb_unpair, (float)m)), lut);
            sM = lse(sM, __fadd_rn(m1, rm), lut);
          } else {
            const float xx = __fadd_rn(r, dev->coeff_num_branches);
            sM1 = lse(sM1, xx, lut);
            sM = lse(sM, __fadd_rn(m1, xx), lut);
          }
        }
        mE[od + i] = sE;
        Mcur[i] = sM;
        sM1 = lse(sM1, sM, lut);
        mM1[od + i] = sM1;
      }
      X::sync();
    }

    // ================================ outside ======================================================
    // keep sums_external[0][*] and [*][L-1], then recycle: E -> P, R -> PM, X -> PM2
    for (int x = c0; x < L; x += cs) {
      E0[x] = X::ld(&mE[X::off(x, L)]);
      EL[x] = X::ld(&mE[X::off(L - 1 - x, L) + x]);
    }
    X::sync();
    const float Z = X::ld(&E0[L - 1]);
    for (ofs_t x = c0; x < TRI; x += cs) { mE[x] = NEG; mR[x] = NEG; mX[x] = NEG; }
    if (c0 == 0 && a.out_logz) a.out_logz[sidx] = Z;
    X::sync();

    const int d_out0 = CONTRA ? (a.allows_short ? 1 : MINSPAN - 1) : (MINSPAN - 1);
    for (int d = L - 1; d >= d_out0; d--) {
      const int ncell = L - d;
      const ofs_t od = X::off(d, L);
      for (int i = c0; i < ncell; i += cs) {
        const int j = i + d;
        // (1) probs_multibranch / probs_multibranch2  (src/mccaskill_algo.rs:540-557, 641-661)
        float pm = NEG, pm2 = NEG;
        const int mmax = L - 1 - d;   // largest k - j over the diagonal (lane i = 0)
        for (int m = 1; m <= mmax; m++) {
          const int k = j + m;
          const bool inr = k < L;
          const ofs_t q = X::off(d + m, L) + i;
          const float c = inr ? X::ld(&mC[q]) : NEG;
          const bool on = c > NEG;
          if (__any_sync(__activemask(), on)) {
            const int kk = inr ? k : j;
            const float p = inr ? X::ld(&mE[q]) : NEG;
            const float x = __fsub_rn(__fadd_rn(p, m_mbclose<CONTRA>(T, s, L, i, kk)), c);
            const float m1 = (m >= 2 && inr) ? X::ld(&mM1[X::off(m - 2, L) + j + 1]) : NEG;
            pm = lse_if(on, pm, __fadd_rn(x, m1), lut);
            if constexpr (CONTRA) pm2 = lse_if(on, pm2, __fadd_rn(x, __fmul_rn(dev->mb_unpair, (float)(m - 1))), lut);
            else pm2 = lse_if(on, pm2, x, lut);
          }
        }
        mR[od + i] = pm;
        mX[od + i] = pm2;
        // (2) the pair (i,j) itself
        const float Cij = X::ld(&mC[od + i]);
        const bool has = Cij > NEG;
        if (__any_sync(__activemask(), has)) {
          const float Aij = has ? __fadd_rn(Cij, m_acc<CONTRA>(T, s, L, i, j)) : NEG;
          const float El = (i < 1) ? 0.f : X::ld(&E0[i - 1]);
          const float Er = (j > L - 2) ? 0.f : X::ld(&EL[j + 1]);
          float sm;
          if constexpr (CONTRA) {
            sm = __fsub_rn(__fadd_rn(__fadd_rn(__fadd_rn(El, Er), Aij), dev->ext_bp), Z);
          } else {
            sm = __fsub_rn(__fadd_rn(__fadd_rn(El, Aij), Er), Z);
          }
          if (!has) sm = NEG;
          // enclosing two-loops: k descending from i-1, l ascending from j+1
          const int cap = min(MAX2, L - d - 3);
          for (int aa = 0; aa <= cap; aa++) {
            const int k = i - 1 - aa;
            for (int bb = 0; aa + bb <= cap; bb++) {
              const int l = j + 1 + bb;
              const bool inr = (k >= 0) && (l < L);
              const ofs_t q = X::off(d + 2 + aa + bb, L) + k;
              const float c = inr ? X::ld(&mC[q]) : NEG;
              const bool on = has && (c > NEG);
              if (__any_sync(__activemask(), on)) {
                const int kk = inr ? k : i, ll = inr ? l : j;
                const float p = inr ? X::ld(&mE[q]) : NEG;
                const float tl = on ? m_twoloop<CONTRA>(T, s, kk, ll, i, j, aa, bb) : 0.f;
                const float y = __fadd_rn(__fsub_rn(__fadd_rn(p, Cij), c), tl);
                sm = lse_if(on, sm, y, lut);
              }
            }
          }
          // enclosing multiloops: k ascending 0..i-1  <=>  m = i-1-k descending
          float sa;
          if constexpr (CONTRA) sa = __fadd_rn(Aij, dev->mb_bp); else sa = __fadd_rn(Aij, dev->coeff_num_branches);
          for (int m = ncell - 2; m >= 0; m--) {
            const bool on = has && (m <= i - 1);
            const int k = on ? (i - 1 - m) : 0;
            const ofs_t q = X::off(d + 1 + m, L) + k;
            const float x = (on && m >= 1) ? X::ld(&mM1[X::off(m - 1, L) + k + 1]) : NEG;
            const float p2 = on ? X::ld(&mX[q]) : NEG;
            const float y = on ? X::ld(&mR[q]) : NEG;
            sm = lse_if(on, sm, __fadd_rn(__fadd_rn(sa, p2), x), lut);
            if constexpr (CONTRA) sm = lse_if(on, sm, __fadd_rn(__fadd_rn(sa, y), __fmul_rn(dev->mb_unpair, (float)m)), lut);
            else sm = lse_if(on, sm, __fadd_rn(sa, y), lut);
            sm = lse_if(on, sm, __fadd_rn(__fadd_rn(sa, x), y), lut);
          }
          if (has && sm > NEG) mE[od + i] = sm;
        }
      }
      X::sync();
    }

    // ================================ BPP = expf(P) ================================================
    for (ofs_t x = c0; x < TRI; x += cs) {
      const float v = X::ld(&mE[x]);
      mE[x] = (v > NEG) ? approx_expf(v) : -1.0f;
    }
    X::sync();
    if (a.out_bpp) {
      float* ob = a.out_bpp + a.bpp_offsets[sidx];
      for (int i = 0; i < L - 1; i++) {
        const size_t rowoff = (size_t)i * (size_t)(2 * L - i - 1) / 2;
        for (int x = c0; x < L - 1 - i; x += cs) ob[rowoff + x] = X::ld(&mE[X::off(x + 1, L) + i]);
      }
    }

    // ================================ centroid (src/centroid_fold.rs:25-105) =======================
    {
      const float* P = mE;
      auto getp = [=](int d, int i) -> float { return X::ld(&P[X::off(d, L) + i]); };
      centroid_run<MODE>(a, sidx, sbeg, L, mR, tstack, getp);
    }
  }
}

__device__ __forceinline__ int coop_doff(int d, int L) { return d * L - ((d * (d - 1)) >> 1); }   // diagonal-major offset
// Grid-wide barrier of the cooperative kernel: one arrival per CTA on a counter that only grows (zeroed by the host
// before the launch), release/acquire at GPU scope by the CTA's first thread, CTA barriers on both sides.
__device__ __forceinline__ void coop_barrier(unsigned* ctr, unsigned& target) {
  target += gridDim.x;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    unsigned seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(ctr) : "memory");
    } while ((int)(seen - target) < 0);
  }
  __syncthreads();
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Max-plus fill of centroid_fold (src/centroid_fold.rs:33-64) for ONE long sequence on the whole grid.  A warp task is
// 32 neighbouring cells of the diagonal (coalesced loads) times a chunk of CENT_CHUNK split points; the partial maxima
// meet in an integer atomicMax: every W is >= +0 and W starts at 0, so the order of the float bits as signed integers
// is the float order wherever it matters, and a maximum does not depend on the order of its operands.
#define RNA_CENT_CHUNK 64
// All thresholds are filled in the same sweep over the diagonals (tasks = threshold x cell group x chunk): the number
// of grid barriers does not grow with the number of thresholds (the reference's default is a sweep of 18).
template <class GETP, class SYNC>
__device__ __forceinline__ void centroid_fill_coop(float* Wall, size_t wstride, int L, const float* gammas, int ngam, GETP getp,
                                                   SYNC sync, int gw, int nw, int lane32) {
  for (int d = 1; d < L; d++) {
    const int ncell = L - d, ng = (ncell + 31) >> 5, nch = max(1, (d - 1 + RNA_CENT_CHUNK - 1) / RNA_CENT_CHUNK);
    const int od = coop_doff(d, L), od1 = coop_doff(d - 1, L);
    for (int tau = gw; tau < ngam * ng * nch; tau += nw) {
      const int gi = tau / (ng * nch), rest = tau - gi * (ng * nch);
      const int g = rest % ng, ch = rest / ng, i = g * 32 + lane32;
      if (i >= ncell) continue;
      float* W = Wall + (size_t)gi * wstride;
      const float gamma = gammas[gi];
      float wv = 0.f;
      if (ch == 0) {
        wv = __ldcg(&W[od1 + i + 1]);
        float e = __ldcg(&W[od1 + i]);
        if (e > wv) wv = e;
        const float p = getp(d, i);
        if (p != -1.0f) {
          const float inner = (d >= 2) ? __ldcg(&W[coop_doff(d - 2, L) + i + 1]) : 0.f;
          e = __fsub_rn(__fadd_rn(inner, __fmul_rn(gamma, p)), 1.0f);
          if (e > wv) wv = e;
        }
      }
      const int m_lo = 1 + ch * RNA_CENT_CHUNK, m_hi = min(d, m_lo + RNA_CENT_CHUNK);
      const float* pa = W + (coop_doff(m_lo, L) + i);                      // W[i][i+m]
      const float* pb = W + (coop_doff(d - m_lo - 1, L) + i + m_lo + 1);   // W[i+m+1][j]
      int sa = L - m_lo, sb = d - m_lo - L - 1;
#pragma unroll 8
      for (int m = m_lo; m < m_hi; m++) {
        const float e = __fadd_rn(__ldcg(pa), __ldcg(pb));
        if (e > wv) wv = e;
        pa += sa; sa--;     // coop_doff(m+1) - coop_doff(m) = L - m
        pb += sb; sb--;     // coop_doff(d-m-2) + i+m+2 - (coop_doff(d-m-1) + i+m+1) = d - m - L - 1
      }
      atomicMax(reinterpret_cast<int*>(&W[od + i]), __float_as_int(wv));
    }
    sync();
  }
}

// centroid_fold for every threshold of the call, one long sequence on the whole grid: zero, one fill sweep, then the
// traceback of threshold g by the first warp of CTA g (mod grid).  Wall: n_gammas matrices wstride floats apart;
// tstacks: n_gammas stacks of 2 (L + 2) ints.
template <class PF, class SYNC>
__device__ __forceinline__ void centroid_coop_all(const FoldArgs& a, uint32_t sidx, uint32_t sbeg, int L, float* Wall,
                                                  size_t wstride, int* tstacks, PF getp, SYNC sync) {
  const int tid = threadIdx.x;
  const size_t gtid = (size_t)blockIdx.x * blockDim.x + tid, gnt = (size_t)gridDim.x * blockDim.x;
  const size_t TRI = (size_t)L * (L + 1) / 2;
  const int ngam = (int)a.n_gammas;
  if (ngam == 0) return;
  for (int g = 0; g < ngam; g++) {
    float* W = Wall + (size_t)g * wstride;
    for (size_t x = gtid; x < TRI; x += gnt) W[x] = 0.f;
    if (a.out_structs) {
      uint8_t* ostr = a.out_structs + (size_t)g * a.total_len + sbeg;
      for (size_t x = gtid; x < (size_t)L; x += gnt) ostr[x] = '.';
    }
  }
  sync();
  const int gw = (tid >> 5) * (int)gridDim.x + (int)blockIdx.x;
  centroid_fill_coop(Wall, wstride, L, a.gammas, ngam, getp, sync, gw, (int)(gridDim.x * (blockDim.x >> 5)), tid & 31);
  if (tid < 32)
    for (int g = (int)blockIdx.x; g < ngam; g += (int)gridDim.x)
      centroid_traceback<MODE_COOP>(a, sidx, sbeg, L, Wall + (size_t)g * wstride, tstacks + (size_t)g * 2 * (L + 2), getp, (uint32_t)g);
  sync();
}
// centroid_fold over packed BPP matrices that already live in device memory (rna_centroid_batch).
// Workspace layout per slot: W (L(L+1)/2 floats) | traceback stack.
template <int MODE>
__global__ void __launch_bounds__(MODE == MODE_SMEM ? 256 : 512) centroid_kernel(const FoldArgs a, const float* bpp_in) {
  typedef Ctx<MODE> X;
  typedef typename X::ofs_t ofs_t;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ int s_work;
  const int tid = threadIdx.x;
  unsigned bar_target = 0;
  (void)bar_target;
  for (uint32_t wloop = 0;; wloop++) {
    uint32_t w;
    if (MODE == MODE_COOP) {
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
    const ofs_t TRI = (ofs_t)L * (ofs_t)(L + 1) / 2;
    float* W;
    if (MODE == MODE_SMEM) W = reinterpret_cast<float*>(smem_raw);
    else if (MODE == MODE_GLOBAL) W = a.workspace + (size_t)blockIdx.x * a.ws_stride;
    else W = a.workspace;
    int* tstack = reinterpret_cast<int*>(W + TRI);
    const float* P = bpp_in + a.bpp_offsets[sidx];
    auto getp = [=](int d, int i) -> float {
      return __ldg(&P[(size_t)i * (size_t)(2 * L - i - 1) / 2 + (size_t)(d - 1)]);
    };
    if constexpr (MODE == MODE_COOP) {
      // one long sequence on the whole grid: all thresholds in one chunked atomic-max fill sweep, own grid barrier
      // (a.work_counter is free here); workspace = n_gammas x (W | traceback stack)
      unsigned* const bar_ctr = reinterpret_cast<unsigned*>(a.work_counter);
      auto grid_sync = [&]() { coop_barrier(bar_ctr, bar_target); };
      const size_t wstride = (size_t)TRI;
      int* tstacks = reinterpret_cast<int*>(a.workspace + (size_t)a.n_gammas * wstride);
      centroid_coop_all(a, sidx, sbeg, L, a.workspace, wstride, tstacks, getp, grid_sync);
    } else {
      centroid_run<MODE>(a, sidx, sbeg, L, W, tstack, getp);
    }
  }
}
__host__ __device__ inline size_t centroid_ws_floats(int L) {
  return (size_t)L * ((size_t)L + 1) / 2 + 2 * ((size_t)L + 2) + 8;
}

}  // namespace rna
