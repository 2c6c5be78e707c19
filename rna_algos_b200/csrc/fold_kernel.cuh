// fold_kernel.cuh — what the fold kernels share: launch arguments, storage modes, the grid-wide barrier of the
// cooperative kernel, and centroid_fold (src/centroid_fold.rs:25-105: max-plus fill + traceback) both fused behind
// the McCaskill passes (fold_kernel2.cuh) and as kernels of its own over packed BPP matrices (rna_centroid_batch).
//
// Matrices are DIAGONAL-MAJOR (index(i,j) = off(j-i) + i): the cells of one diagonal are contiguous, so the 32 lanes
// of a warp (cells i..i+31 of one diagonal) touch 32 consecutive words for every operand of every recurrence.
//   MODE_SMEM   one CTA per sequence, working set in shared memory
//   MODE_GLOBAL one CTA per sequence, working set in an HBM/L2 workspace slot
//   MODE_COOP   the whole grid works on one sequence, grid-wide barrier per diagonal
// (The first kernel of round 1, a thread-per-cell wavefront over these modes, is gone from the product: see git
// history, commit ed704ea.)
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
  float* sums;               // optional FoldSums / FoldScores planes (rna_fold_sums_batch), else null
  const unsigned long long* sums_offsets;
  int inside_only;           // stop after the inside pass
  unsigned ring_bytes;          // HBM-resident mode: bytes of the dense chains' shared-memory operand ring (0 = none)
  int nXw_out;                  // warps of role X in the outside pass (0 = as in the inside pass); the rest fold role Y
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
