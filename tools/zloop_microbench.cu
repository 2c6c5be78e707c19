// zloop_microbench.cu — does register prefetching hide L2 latency under the three-chain fold of role Z? (dev aid)
#include <cstdio>
#include <cuda_runtime.h>
#include "../rna_algos_b200/csrc/numerics.cuh"
using namespace rna;

template <int PF, bool CG>
__global__ void zloop(const float* __restrict__ R, const float* __restrict__ E, const float* __restrict__ M1, int L, int d,
                      float* out, long long* cyc, float coeff) {
  extern __shared__ float4 lut[];
  load_lse_lut(lut);
  __syncthreads();
  const int i = threadIdx.x;
  auto off = [&](int dd) { return dd * L - ((dd * (dd - 1)) >> 1); };
  auto ld = [&](const float* p) { return CG ? __ldcg(p) : *p; };
  float sE = 0.f, sM1 = -1.f, sM = RNA_NEG_INF;
  float pr[PF], pe[PF], pq[PF];
#pragma unroll
  for (int u = 0; u < PF; u++) { const int m = 1 + u; pr[u] = ld(&R[off(d - m) + i + m]); pe[u] = ld(&E[off(m - 1) + i]); pq[u] = ld(&M1[off(m - 1) + i]); }
  const long long t0 = clock64();
  for (int m0 = 1; m0 < d; m0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const int m = m0 + u;
      if (m < d) {
        const float r = pr[u], e = pe[u], m1 = pq[u];
        const int mn = m + PF;
        if (mn < d) { pr[u] = ld(&R[off(d - mn) + i + mn]); pe[u] = ld(&E[off(mn - 1) + i]); pq[u] = ld(&M1[off(mn - 1) + i]); }
        sE = lse(sE, __fadd_rn(r, e), lut);
        const float xx = __fadd_rn(r, coeff);
        sM1 = lse(sM1, xx, lut);
        sM = lse(sM, __fadd_rn(m1, xx), lut);
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = sE + sM1 + sM;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int PF, bool CG>
void run(const float* R, const float* E, const float* M1, int L, int d, int blocks, int threads) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * blocks * threads); cudaMalloc(&cyc, 8 * blocks);
  zloop<PF, CG><<<blocks, threads, 256>>>(R, E, M1, L, d, out, cyc, -1.5f);
  zloop<PF, CG><<<blocks, threads, 256>>>(R, E, M1, L, d, out, cyc, -1.5f);
  cudaDeviceSynchronize();
  long long h[256];
  cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
  printf("PF=%d %s blocks=%d threads=%d: %.1f cycles per split point\n", PF, CG ? "ld.cg" : "ld", blocks, threads, (double)h[0] / (d - 1));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  const int L = 2048, d = 1500;
  const size_t T = (size_t)L * (L + 1) / 2;
  float *R, *E, *M1;
  cudaMalloc(&R, 4 * T); cudaMalloc(&E, 4 * T); cudaMalloc(&M1, 4 * T);
  float* h = (float*)malloc(4 * T);
  for (size_t x = 0; x < T; x++) h[x] = -0.001f * (float)(x % 7919);
  cudaMemcpy(R, h, 4 * T, cudaMemcpyHostToDevice); cudaMemcpy(E, h, 4 * T, cudaMemcpyHostToDevice); cudaMemcpy(M1, h, 4 * T, cudaMemcpyHostToDevice);
  run<1, false>(R, E, M1, L, d, 1, 32);
  run<2, false>(R, E, M1, L, d, 1, 32);
  run<4, false>(R, E, M1, L, d, 1, 32);
  run<6, false>(R, E, M1, L, d, 1, 32);
  run<6, true>(R, E, M1, L, d, 1, 32);
  run<8, false>(R, E, M1, L, d, 1, 32);
  run<6, false>(R, E, M1, L, d, 17, 32);     // 17 warps on different SMs = the 544 cells of the diagonal
  run<6, false>(R, E, M1, L, d, 148, 128);
  return 0;
}
