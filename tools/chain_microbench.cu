// chain_microbench.cu — cycles per split point of ONE dense chain per lane (the cooperative kernel's role Z),
// operands streamed from an L2/HBM-resident triangular matrix (development aid).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -fmad=false -o chain_microbench chain_microbench.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../rna_algos_b200/csrc/fold_phases.cuh"
using namespace rna;

// variant 1: branch-free folds (operands of m >= d are -inf: no-ops), running pointers instead of 64-bit index math
template <int PF, bool HASB, class OP>
__device__ __forceinline__ float chain_fold_v1(const float* __restrict__ A, const float* __restrict__ B, int L, int d, int i,
                                               float sum, const float4* lut, OP op) {
  const float NEG = RNA_NEG_INF;
  const float* pA = A + doff(d - 1, L) + i + 1;
  const float* pB = B + i;
  int sA = d - 1 - L, sB = L, mL = 1;
  float pa[PF], pb[PF];
  auto load = [&](float& a, float& b) {
    a = NEG; b = NEG;
    if (mL < d) { a = *pA; if (HASB) b = *pB; }
    pA += sA; pB += sB; sA--; sB--; mL++;
  };
#pragma unroll
  for (int u = 0; u < PF; u++) load(pa[u], pb[u]);
  for (int m0 = 1; m0 < d; m0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const float a = pa[u], b = pb[u];
      load(pa[u], pb[u]);
      sum = lse(sum, op(m0 + u, a, b), lut);
    }
  }
  return sum;
}
// variant 3: variant 1 with the latency-optimised logsumexp
template <int PF, bool HASB, class OP>
__device__ __forceinline__ float chain_fold_v3(const float* __restrict__ A, const float* __restrict__ B, int L, int d, int i,
                                               float sum, OP op) {
  const float NEG = RNA_NEG_INF;
  const float* pA = A + doff(d - 1, L) + i + 1;
  const float* pB = B + i;
  int sA = d - 1 - L, sB = L, mL = 1;
  float pa[PF], pb[PF];
  auto load = [&](float& a, float& b) {
    a = NEG; b = NEG;
    if (mL < d) { a = *pA; if (HASB) b = *pB; }
    pA += sA; pB += sB; sA--; sB--; mL++;
  };
#pragma unroll
  for (int u = 0; u < PF; u++) load(pa[u], pb[u]);
  for (int m0 = 1; m0 < d; m0 += PF) {
#pragma unroll
    for (int u = 0; u < PF; u++) {
      const float a = pa[u], b = pb[u];
      load(pa[u], pb[u]);
      sum = lse_lat(sum, op(m0 + u, a, b));
    }
  }
  return sum;
}
// variant 2: no memory at all (pure logsumexp chain floor)
__device__ __forceinline__ float chain_nomem(int d, float sum, float x, const float4* lut) {
  for (int m = 1; m < d; m++) { sum = lse(sum, x, lut); x = __fadd_rn(x, 0.37f); if (x > 3.f) x = __fadd_rn(x, -9.f); }
  return sum;
}

template <int VAR, int PF>
__global__ void bench(const float* A, const float* B, float* out, long long* cyc, int L, int d, int warps_used) {
  __shared__ float4 lut[8];
  load_lse_lut(lut);
  __syncthreads();
  const int w = (threadIdx.x >> 5) * gridDim.x + blockIdx.x;   // consecutive warps on different SMs
  if (w >= warps_used) return;
  const int i = (w * 32 + (threadIdx.x & 31)) % (L - d);
  const long long t0 = clock64();
  float r;
  auto op = [](int, float a, float b) { return __fadd_rn(a, b); };
  if (VAR == 0) r = chain_fold<PF, true>(A, B, L, d, i, -3.f, lut, op);
  else if (VAR == 1) r = chain_fold_v1<PF, true>(A, B, L, d, i, -3.f, lut, op);
  else if (VAR == 3) r = chain_fold_v3<PF, true>(A, B, L, d, i, -3.f, op);
  else r = chain_nomem(d, -3.f, -1.f - 0.01f * i, lut);
  const long long t1 = clock64();
  out[w * 32 + (threadIdx.x & 31)] = r;
  if ((threadIdx.x & 31) == 0) cyc[w] = t1 - t0;
}

template <int VAR, int PF>
void run(const char* name, const float* A, const float* B, int L, int d, int warps) {
  float* out; long long* cyc;
  const int grid = 148, nt = 256;
  cudaMalloc(&out, 4 * grid * nt); cudaMalloc(&cyc, 8 * grid * 8);
  for (int rep = 0; rep < 2; rep++) bench<VAR, PF><<<grid, nt>>>(A, B, out, cyc, L, d, warps);
  cudaDeviceSynchronize();
  std::vector<long long> h(warps);
  cudaMemcpy(h.data(), cyc, 8 * warps, cudaMemcpyDeviceToHost);
  long long mn = h[0], mx = h[0]; double avg = 0;
  for (auto v : h) { mn = std::min(mn, v); mx = std::max(mx, v); avg += v; }
  printf("%-34s L=%d d=%d warps=%4d: cycles per split point min %.0f avg %.0f max %.0f   (%s)\n", name, L, d, warps,
         (double)mn / (d - 1), avg / warps / (d - 1), (double)mx / (d - 1), cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int L : {2048}) {
    const size_t TRI = (size_t)L * (L + 1) / 2;
    std::vector<float> h(TRI);
    unsigned st = 1234567u;
    for (auto& x : h) { st = st * 1664525u + 1013904223u; x = -20.f * ((st >> 8) * (1.f / 16777216.f)); }
    float *A, *B;
    cudaMalloc(&A, TRI * 4); cudaMalloc(&B, TRI * 4);
    cudaMemcpy(A, h.data(), TRI * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(B, h.data(), TRI * 4, cudaMemcpyHostToDevice);
    const int d = L * 3 / 4;
    for (int warps : {148}) {
      run<2, 1>("no memory (lse floor)", A, B, L, d, warps);
      run<0, 8>("chain_fold PF=8 (branchy)", A, B, L, d, warps);
      run<3, 2>("v3 lse_lat PF=2", A, B, L, d, warps);
      run<3, 4>("v3 lse_lat PF=4", A, B, L, d, warps);
      run<3, 8>("v3 lse_lat PF=8", A, B, L, d, warps);
      run<1, 4>("v1 branch-free PF=4", A, B, L, d, warps);
      run<1, 8>("v1 branch-free PF=8", A, B, L, d, warps);
      run<1, 16>("v1 branch-free PF=16", A, B, L, d, warps);
    }
    cudaFree(A); cudaFree(B);
  }
  return 0;
}
