// durbin_kernel.cuh — Durbin's 3-state pair-HMM forward-backward as a 2-D anti-diagonal wavefront.
//
// Reference: src/durbin_algo.rs:79-199 (get_align_sums), :201-242 (get_match_probs).  One thread per
// row i of the current anti-diagonal s = i + j.  Only three anti-diagonals of each state are live, so
// F_I/F_D/B_M/B_I/B_D roll through 3 x n buffers; F_M (needed by the posterior) is parked in an n x m matrix
// of the CTA's own (shared memory when it fits, else a global slot) and overwritten there by the match probability
// when the backward sweep reaches (i+1, j+1).  The posterior is fused into the backward sweep.  The output matrix is
// written ONCE at the end, whole rows at a time (write-only and coalesced: it may be the caller's page-locked host
// buffer, so the device-to-host transfer runs under the kernel).  All folds are in the reference's order, so the
// result is bit-identical to the reference.
#pragma once
#include "dev_tables.h"
#include "numerics.cuh"

namespace rna {

struct DurbinArgs {
  const uint8_t* bases;
  const uint32_t* offsets;
  const uint32_t* pairs;
  const uint32_t* order;       // pair indices in launch order, or null
  uint32_t n_pairs;
  const uint64_t* prob_offsets;
  float* out_probs;
  const DevAlign* tables;
  float* workspace;            // used when the rolling buffers do not fit in shared memory
  unsigned long long ws_stride;
  int* work_counter;
  int ncap, mcap;              // shared-memory capacity (padded lengths)
  int roll_in_smem;
  int park_in_smem;            // the n x m forward / posterior matrix lives in shared memory
  float* park_ws;              // else: one slot of park_stride floats per CTA
  unsigned long long park_stride;
};

__host__ __device__ inline size_t durbin_smem_bytes(int ncap, int mcap, bool roll_in_smem, bool park_in_smem = false) {
  size_t b = 128 + 128;   // LSE LUT + DevAlign copy
  b += ((size_t)ncap + (size_t)mcap + 8 + 15) / 16 * 16;
  if (roll_in_smem) b += (size_t)9 * (size_t)ncap * 4;
  if (park_in_smem) b += (size_t)ncap * (size_t)mcap * 4;
  return b;
}

__global__ void __launch_bounds__(256, 6) durbin_kernel(const DurbinArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* lut = reinterpret_cast<float4*>(smem_raw);
  DevAlign* T = reinterpret_cast<DevAlign*>(smem_raw + 128);
  uint8_t* sseq = smem_raw + 256;
  float* sroll = reinterpret_cast<float*>(smem_raw + 256 + ((size_t)a.ncap + (size_t)a.mcap + 8 + 15) / 16 * 16);
  __shared__ int s_work;
  const int tid = threadIdx.x, nt = blockDim.x;
  load_lse_lut(lut);
  if (tid < (int)(sizeof(DevAlign) / 4)) reinterpret_cast<float*>(T)[tid] = reinterpret_cast<const float*>(a.tables)[tid];
  __syncthreads();
  const float NEG = RNA_NEG_INF;
  const float m2m = T->m2m, m2i = T->m2i, iex = T->iex, inm = T->inm, ini = T->ini;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_work = atomicAdd(a.work_counter, 1);
    __syncthreads();
    const uint32_t w = (uint32_t)s_work;
    if (w >= a.n_pairs) break;
    const uint32_t pidx = a.order ? a.order[w] : w;
    const uint32_t sa = a.pairs[2 * pidx], sb = a.pairs[2 * pidx + 1];
    const uint32_t ba = a.offsets[sa], bb = a.offsets[sb];
    const int n = (int)(a.offsets[sa + 1] - ba) + 2, m = (int)(a.offsets[sb + 1] - bb) + 2;
    uint8_t* p0 = sseq;            // sentinel-padded copies (src/bin/durbin_algo.rs:48-50)
    uint8_t* p1 = sseq + n;
    for (int x = tid; x < n; x += nt) p0[x] = (x == 0 || x == n - 1) ? 4 : a.bases[ba + x - 1];
    for (int x = tid; x < m; x += nt) p1[x] = (x == 0 || x == m - 1) ? 4 : a.bases[bb + x - 1];
    float* roll = a.roll_in_smem ? sroll : a.workspace + (size_t)blockIdx.x * a.ws_stride;
    float* RM = roll;               // [3][n]
    float* RI = roll + 3 * (size_t)n;
    float* RD = roll + 6 * (size_t)n;
    float* out = a.out_probs + a.prob_offsets[pidx];
    float* park = a.park_in_smem ? sroll + (a.roll_in_smem ? (size_t)9 * a.ncap : 0)
                                 : a.park_ws + (size_t)blockIdx.x * a.park_stride;
    for (int x = tid; x < 9 * n; x += nt) roll[x] = NEG;
    __syncthreads();

    // ---------------- forward: src/durbin_algo.rs:82-139 ------------------------------------------
    for (int s = 0; s <= n + m - 4; s++) {
      const int b0 = (s % 3) * n, b1 = ((s + 2) % 3) * n, b2 = ((s + 1) % 3) * n;   // s, s-1, s-2
      const int ilo = max(0, s - (m - 2)), ihi = min(n - 2, s);
      for (int i = ilo + tid; i <= ihi; i += nt) {
        const int j = s - i;
        float fm = NEG, fi = NEG, fd = NEG;
        if (i == 0 && j == 0) {
          fm = 0.f;
        } else {
          if (i > 0 && j > 0) {
            const bool begins = (i == 1 && j == 1);
            float sum = lse_init(__fadd_rn(RM[b2 + i - 1], begins ? inm : m2m));
            sum = lse(sum, __fadd_rn(RI[b2 + i - 1], m2i), lut);
            sum = lse(sum, __fadd_rn(RD[b2 + i - 1], m2i), lut);
            fm = __fadd_rn(sum, T->match[p0[i] * 4 + p1[j]]);
          }
          if (i > 0) {
            const bool begins = (i == 1 && j == 0);
            float sum = lse_init(__fadd_rn(RM[b1 + i - 1], begins ? ini : m2i));
            sum = lse(sum, __fadd_rn(RI[b1 + i - 1], iex), lut);
            fi = __fadd_rn(sum, T->insert[p0[i]]);
          }
          if (j > 0) {
            const bool begins = (i == 0 && j == 1);
            float sum = lse_init(__fadd_rn(RM[b1 + i], begins ? ini : m2i));
            sum = lse(sum, __fadd_rn(RD[b1 + i], iex), lut);
            fd = __fadd_rn(sum, T->insert[p1[j]]);
          }
        }
        RM[b0 + i] = fm; RI[b0 + i] = fi; RD[b0 + i] = fd;
        park[(size_t)i * m + j] = fm;     // park forward_sums_match for the posterior
      }
      __syncthreads();
    }
    // global_sum: src/durbin_algo.rs:207-215 (fold starts from F_M)
    float Z;
    {
      const int b0 = ((n + m - 4) % 3) * n;
      Z = RM[b0 + n - 2];
      Z = lse(Z, RI[b0 + n - 2], lut);
      Z = lse(Z, RD[b0 + n - 2], lut);
    }
    __syncthreads();
    for (int x = tid; x < 9 * n; x += nt) roll[x] = NEG;
    __syncthreads();

    // ---------------- backward + posterior: src/durbin_algo.rs:140-197, 216-240 -------------------
    for (int s = n + m - 2; s >= 2; s--) {
      const int b0 = (s % 3) * n, b1 = ((s + 1) % 3) * n, b2 = ((s + 2) % 3) * n;   // s, s+1, s+2
      const int ilo = max(1, s - (m - 1)), ihi = min(n - 1, s - 1);
      for (int i = ilo + tid; i <= ihi; i += nt) {
        const int j = s - i;
        float bm = NEG, bi = NEG, bd = NEG;
        const bool corner = (i == n - 1 && j == m - 1);
        if (corner) {
          bm = 0.f;
        } else {
          if (i < n - 1 && j < m - 1) {
            const bool ends = (i + 1 == n - 1 && j + 1 == m - 1);
            float sum = lse_init(__fadd_rn(RM[b2 + i + 1], ends ? 0.f : m2m));
            sum = lse(sum, __fadd_rn(RI[b2 + i + 1], m2i), lut);
            sum = lse(sum, __fadd_rn(RD[b2 + i + 1], m2i), lut);
            bm = __fadd_rn(sum, T->match[p0[i] * 4 + p1[j]]);
          }
          if (i < n - 1) {
            const bool ends = (i + 1 == n - 1 && j == m - 1);
            float sum = lse_init(__fadd_rn(RM[b1 + i + 1], ends ? 0.f : m2i));
            sum = lse(sum, __fadd_rn(RI[b1 + i + 1], iex), lut);
            bi = __fadd_rn(sum, T->insert[p0[i]]);
          }
          if (j < m - 1) {
            const bool ends = (i == n - 1 && j + 1 == m - 1);
            float sum = lse_init(__fadd_rn(RM[b1 + i], ends ? 0.f : m2i));
            sum = lse(sum, __fadd_rn(RD[b1 + i], iex), lut);
            bd = __fadd_rn(sum, T->insert[p1[j]]);
          }
        }
        RM[b0 + i] = bm; RI[b0 + i] = bi; RD[b0 + i] = bd;
        if (i >= 2 && j >= 2) {
          // match probability of (i-1, j-1): uses backward sums at (i, j)
          float t = lse_init(__fadd_rn(corner ? 0.f : m2m, bm));
          t = lse(t, __fadd_rn(m2i, bi), lut);
          t = lse(t, __fadd_rn(m2i, bd), lut);
          const size_t q = (size_t)(i - 1) * m + (j - 1);
          const float fwd = park[q];
          park[q] = approx_expf(__fsub_rn(__fadd_rn(fwd, t), Z));
        }
      }
      __syncthreads();
    }
    // the output, row-major n x m with a zero border (rows 0, n-1; columns 0, m-1), in one coalesced sweep
    // (64-bit cell count: n * m exceeds 2^31 when both sequences are longer than 46340 nt)
    const size_t nm = (size_t)n * (size_t)m;
    for (size_t x = tid; x < nm; x += nt) {
      const int i = (int)(x / (unsigned)m), j = (int)(x - (size_t)i * m);
      out[x] = (i == 0 || i == n - 1 || j == 0 || j == m - 1) ? 0.f : park[x];
    }
  }
}

}  // namespace rna
