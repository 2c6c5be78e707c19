// lse_microbench.cu — cycles per dependent reference-exact logsumexp step on one warp (development aid).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -o lse_microbench lse_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../rna_algos_b200/csrc/numerics.cuh"
using namespace rna;

// variant 1: 9-row LUT (row 8 = identity for z >= threshold), finite handling folded into one select
__device__ __constant__ float4 kCoef9[9] = {
    {-0.0065591595f, 0.12764427f, 0.49965546f, 0.6931542f},  {-0.015515756f, 0.14467756f, 0.48829398f, 0.6958093f},
    {-0.012890925f, 0.13010283f, 0.51503986f, 0.6795586f},   {-0.0072142647f, 0.087754086f, 0.6208708f, 0.5909676f},
    {-0.0031455354f, 0.046722945f, 0.7592532f, 0.43487945f}, {-0.0010110698f, 0.018594341f, 0.88317305f, 0.25236955f},
    {-0.000196278f, 0.0046084408f, 0.9634432f, 0.09831489f}, {-0.0000113994f, 0.0003734731f, 0.9959107f, 0.0149855051f},
    {0.f, 0.f, 1.f, 0.f}};

__device__ __forceinline__ float lse_v1(float sum, float x, const float4* lut9) {
  const float y = fminf(sum, x);
  const float z = __fsub_rn(fmaxf(sum, x), y);
  // z >= 0 (or NaN/inf): float order == integer order of the bits
  const int zb = __float_as_int(z);
  int idx = (zb >= __float_as_int(0.66153675f)) + (zb >= __float_as_int(1.6320158f)) + (zb >= __float_as_int(2.4912589f)) +
            (zb >= __float_as_int(3.37925f)) + (zb >= __float_as_int(4.426169f)) + (zb >= __float_as_int(5.789071f)) +
            (zb >= __float_as_int(7.8162727f)) + (zb >= __float_as_int(11.862479f));
  idx = min(max(idx, 0), 8);
  const float4 c = lut9[idx];
  float r = __fadd_rn(__fmul_rn(c.x, z), c.y);
  r = __fadd_rn(__fmul_rn(r, z), c.z);
  r = __fadd_rn(__fmul_rn(r, z), c.w);
  const float v = __fadd_rn(y, r);
  const bool fx = is_finite(x), fs = is_finite(sum);
  const float alt = fx ? x : sum;       // off the critical path
  return (fx && fs) ? v : alt;
}

// variant 2: no shared-memory LUT — the seven breakpoint compares run side by side right after z, and the coefficients
// come out of select trees (the tree of `a` is the only one on the chain; b, c, d resolve under the Horner steps)
__device__ __forceinline__ float sel8(bool q1, bool q2, bool q3, bool q4, bool q5, bool q6, bool q7, float c0, float c1,
                                      float c2, float c3, float c4, float c5, float c6, float c7) {
  const float lo = q2 ? (q3 ? c3 : c2) : (q1 ? c1 : c0);
  const float hi = q6 ? (q7 ? c7 : c6) : (q5 ? c5 : c4);
  return q4 ? hi : lo;
}
__device__ __forceinline__ float lse_v2(float sum, float x) {
  x = (x > RNA_NEG_INF) ? x : RNA_NEG_INF;
  const float y = fminf(sum, x);
  const float mx = fmaxf(sum, x);
  const float z = __fsub_rn(mx, y);
  const bool q1 = !(z < 0.66153675f), q2 = !(z < 1.6320158f), q3 = !(z < 2.4912589f), q4 = !(z < 3.37925f),
             q5 = !(z < 4.426169f), q6 = !(z < 5.789071f), q7 = !(z < 7.8162727f);
  const float a = sel8(q1, q2, q3, q4, q5, q6, q7, -0.0065591595f, -0.015515756f, -0.012890925f, -0.0072142647f,
                       -0.0031455354f, -0.0010110698f, -0.000196278f, -0.0000113994f);
  const float b = sel8(q1, q2, q3, q4, q5, q6, q7, 0.12764427f, 0.14467756f, 0.13010283f, 0.087754086f, 0.046722945f,
                       0.018594341f, 0.0046084408f, 0.0003734731f);
  const float c = sel8(q1, q2, q3, q4, q5, q6, q7, 0.49965546f, 0.48829398f, 0.51503986f, 0.6208708f, 0.7592532f,
                       0.88317305f, 0.9634432f, 0.9959107f);
  const float d = sel8(q1, q2, q3, q4, q5, q6, q7, 0.6931542f, 0.6958093f, 0.6795586f, 0.5909676f, 0.43487945f,
                       0.25236955f, 0.09831489f, 0.0149855051f);
  float r = __fadd_rn(__fmul_rn(a, z), b);
  r = __fadd_rn(__fmul_rn(r, z), c);
  r = __fadd_rn(__fmul_rn(r, z), d);
  const float v = __fadd_rn(y, (z >= RNA_LSE_THRESHOLD) ? z : r);
  return (z < __int_as_float(0x7f800000)) ? v : mx;
}
// variant 3: shared-memory LUT, index = number of breakpoints <= z from seven independent compares
__device__ __forceinline__ float lse_v3(float sum, float x, const float4* lut) {
  x = (x > RNA_NEG_INF) ? x : RNA_NEG_INF;
  const float y = fminf(sum, x);
  const float mx = fmaxf(sum, x);
  const float z = __fsub_rn(mx, y);
  const int idx = (int)!(z < 0.66153675f) + (int)!(z < 1.6320158f) + (int)!(z < 2.4912589f) + (int)!(z < 3.37925f) +
                  (int)!(z < 4.426169f) + (int)!(z < 5.789071f) + (int)!(z < 7.8162727f);
  const float4 c = lut[idx];
  float r = __fadd_rn(__fmul_rn(c.x, z), c.y);
  r = __fadd_rn(__fmul_rn(r, z), c.z);
  r = __fadd_rn(__fmul_rn(r, z), c.w);
  const float v = __fadd_rn(y, (z >= RNA_LSE_THRESHOLD) ? z : r);
  return (z < __int_as_float(0x7f800000)) ? v : mx;
}

// variant 4: `a` from a select tree, b/c/d from the shared-memory LUT row whose index comes from the same predicates
__device__ __forceinline__ int isel8(bool q1, bool q2, bool q3, bool q4, bool q5, bool q6, bool q7) {
  const int lo = q2 ? (q3 ? 3 : 2) : (q1 ? 1 : 0);
  const int hi = q6 ? (q7 ? 7 : 6) : (q5 ? 5 : 4);
  return q4 ? hi : lo;
}
__device__ __forceinline__ float lse_v4(float sum, float x, const float4* lut) {
  x = (x > RNA_NEG_INF) ? x : RNA_NEG_INF;
  const float y = fminf(sum, x);
  const float mx = fmaxf(sum, x);
  const float z = __fsub_rn(mx, y);
  const bool q1 = !(z < 0.66153675f), q2 = !(z < 1.6320158f), q3 = !(z < 2.4912589f), q4 = !(z < 3.37925f),
             q5 = !(z < 4.426169f), q6 = !(z < 5.789071f), q7 = !(z < 7.8162727f);
  const float a = sel8(q1, q2, q3, q4, q5, q6, q7, -0.0065591595f, -0.015515756f, -0.012890925f, -0.0072142647f,
                       -0.0031455354f, -0.0010110698f, -0.000196278f, -0.0000113994f);
  const float4 c = lut[isel8(q1, q2, q3, q4, q5, q6, q7)];
  float r = __fadd_rn(__fmul_rn(a, z), c.y);
  r = __fadd_rn(__fmul_rn(r, z), c.z);
  r = __fadd_rn(__fmul_rn(r, z), c.w);
  const float v = __fadd_rn(y, (z >= RNA_LSE_THRESHOLD) ? z : r);
  return (z < __int_as_float(0x7f800000)) ? v : mx;
}
// variant 5: a and b from select trees, c/d from the LUT
__device__ __forceinline__ float lse_v5(float sum, float x, const float4* lut) {
  x = (x > RNA_NEG_INF) ? x : RNA_NEG_INF;
  const float y = fminf(sum, x);
  const float mx = fmaxf(sum, x);
  const float z = __fsub_rn(mx, y);
  const bool q1 = !(z < 0.66153675f), q2 = !(z < 1.6320158f), q3 = !(z < 2.4912589f), q4 = !(z < 3.37925f),
             q5 = !(z < 4.426169f), q6 = !(z < 5.789071f), q7 = !(z < 7.8162727f);
  const float a = sel8(q1, q2, q3, q4, q5, q6, q7, -0.0065591595f, -0.015515756f, -0.012890925f, -0.0072142647f,
                       -0.0031455354f, -0.0010110698f, -0.000196278f, -0.0000113994f);
  const float b = sel8(q1, q2, q3, q4, q5, q6, q7, 0.12764427f, 0.14467756f, 0.13010283f, 0.087754086f, 0.046722945f,
                       0.018594341f, 0.0046084408f, 0.0003734731f);
  const float2 c = reinterpret_cast<const float2*>(lut)[2 * isel8(q1, q2, q3, q4, q5, q6, q7) + 1];
  float r = __fadd_rn(__fmul_rn(a, z), b);
  r = __fadd_rn(__fmul_rn(r, z), c.x);
  r = __fadd_rn(__fmul_rn(r, z), c.y);
  const float v = __fadd_rn(y, (z >= RNA_LSE_THRESHOLD) ? z : r);
  return (z < __int_as_float(0x7f800000)) ? v : mx;
}

template <int VAR, int ILP>
__global__ void bench(float* out, long long* cyc, int iters, float x0) {
  extern __shared__ float4 lut[];
  if (threadIdx.x < 9) lut[threadIdx.x] = (VAR == 0) ? kLnExp1pCoef[min((int)threadIdx.x, 7)] : kCoef9[threadIdx.x];
  __syncthreads();
  float s[ILP];
  for (int k = 0; k < ILP; k++) s[k] = -1.0f * threadIdx.x - k;
  float x = x0 + 0.001f * threadIdx.x;
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < ILP; k++) s[k] = (VAR == 0) ? lse(s[k], x, lut) : (VAR == 1) ? lse_v1(s[k], x, lut) : (VAR == 2) ? lse_v2(s[k], x) : (VAR == 3) ? lse_v3(s[k], x, lut) : (VAR == 4) ? lse_v4(s[k], x, lut) : lse_v5(s[k], x, lut);
    x = __fadd_rn(x, 0.37f);
    if (x > 9.f) x = __fadd_rn(x, -12.f);
  }
  const long long t1 = clock64();
  float acc = 0;
  for (int k = 0; k < ILP; k++) acc += s[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// dependent-instruction latency probes
__global__ void probe(float* out, long long* cyc, int iters) {
  float a = threadIdx.x * 0.5f, b = 1.0001f;
  long long t0 = clock64();
  for (int i = 0; i < iters; i++) { a = __fadd_rn(a, b); a = __fadd_rn(a, b); a = __fadd_rn(a, b); a = __fadd_rn(a, b); }
  long long t1 = clock64();
  float c = a;
  long long t2 = clock64();
  for (int i = 0; i < iters; i++) { c = fminf(c, b) ; c = fmaxf(c, a); c = fminf(c, b + 1.f); c = fmaxf(c, a - 1.f); }
  long long t3 = clock64();
  float d = a;
  long long t4 = clock64();
  for (int i = 0; i < iters; i++) { d = (d < b) ? a : c; d = (d < a) ? b : c; d = (d < c) ? a : b; d = (d < b) ? c : a; }
  long long t5 = clock64();
  out[threadIdx.x] = a + c + d;
  if (threadIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = t3 - t2; cyc[2] = t5 - t4; }
}

template <int VAR, int ILP>
void run(const char* name, int warps, int blocks) {
  float* out; long long* cyc;
  cudaMalloc(&out, 4 * 1024 * 64); cudaMalloc(&cyc, 8 * 64);
  const int iters = 20000;
  bench<VAR, ILP><<<blocks, 32 * warps, 256>>>(out, cyc, iters, 0.3f);
  bench<VAR, ILP><<<blocks, 32 * warps, 256>>>(out, cyc, iters, 0.3f);
  long long h[64];
  cudaDeviceSynchronize();
  cudaMemcpy(h, cyc, 8 * blocks, cudaMemcpyDeviceToHost);
  printf("%-28s warps/CTA=%d: %.1f cycles per step (%.1f per LSE)\n", name, warps, (double)h[0] / iters, (double)h[0] / iters / ILP);
  // check v1 == v0 on a few values
  cudaFree(out); cudaFree(cyc);
}

__global__ void check(int* bad) {
  extern __shared__ float4 lut[];
  float4* l8 = lut; float4* l9 = lut + 8;
  if (threadIdx.x < 8) l8[threadIdx.x] = kLnExp1pCoef[threadIdx.x];
  if (threadIdx.x < 9) l9[threadIdx.x] = kCoef9[threadIdx.x];
  __syncthreads();
  int nb = 0;
  unsigned st = 12345u + threadIdx.x * 7919u + blockIdx.x * 104729u;
  for (int i = 0; i < 200000; i++) {
    st = st * 1664525u + 1013904223u; float a = ((st >> 8) * (1.f / 16777216.f) - 0.5f) * 60.f;
    st = st * 1664525u + 1013904223u; float b = ((st >> 8) * (1.f / 16777216.f) - 0.5f) * 60.f;
    if ((i & 63) == 0) a = RNA_NEG_INF;
    if ((i & 127) == 1) b = RNA_NEG_INF;
    if ((i & 255) == 2) b = a;
    if ((i & 255) == 3) b = __fadd_rn(a, 11.862479f);
    const float r0 = lse(a, b, l8), r1 = lse_v1(a, b, l9), r2 = lse_v2(a, b), r3 = lse_v3(a, b, l8), r4 = lse_v4(a, b, l8), r5 = lse_v5(a, b, l8);
    if (__float_as_int(r0) != __float_as_int(r1) || __float_as_int(r0) != __float_as_int(r2) || __float_as_int(r0) != __float_as_int(r3) || __float_as_int(r0) != __float_as_int(r4) || __float_as_int(r0) != __float_as_int(r5)) nb++;
  }
  if (nb) atomicAdd(bad, nb);
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 64);
  probe<<<1, 32>>>(out, cyc, 10000);
  long long h[3];
  cudaMemcpy(h, cyc, 24, cudaMemcpyDeviceToHost);
  printf("dependent FADD: %.2f cyc, FMNMX: %.2f cyc, FSETP+FSEL: %.2f cyc\n", h[0] / 40000.0, h[1] / 40000.0, h[2] / 40000.0);
  int* bad; cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4);
  check<<<64, 128, 17 * 16>>>(bad);
  int hb; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost);
  printf("lse_v1 vs lse mismatches: %d of %d\n", hb, 64 * 128 * 200000);
  run<0, 1>("lse v0 ILP1", 1, 1);
  run<0, 2>("lse v0 ILP2", 1, 1);
  run<0, 3>("lse v0 ILP3", 1, 1);
  run<0, 1>("lse v0 ILP1", 4, 1);
  run<0, 1>("lse v0 ILP1", 8, 1);
  run<0, 1>("lse v0 ILP1 16 warps", 16, 1);
  run<1, 1>("lse v1 ILP1", 1, 1);
  run<1, 2>("lse v1 ILP2", 1, 1);
  run<1, 3>("lse v1 ILP3", 1, 1);
  run<1, 1>("lse v1 ILP1", 4, 1);
  run<1, 1>("lse v1 ILP1 16 warps", 16, 1);
  run<2, 1>("lse v2 (select trees) ILP1", 1, 1);
  run<2, 2>("lse v2 ILP2", 1, 1);
  run<2, 3>("lse v2 ILP3", 1, 1);
  run<2, 1>("lse v2 ILP1", 4, 1);
  run<2, 1>("lse v2 ILP1 16 warps", 16, 1);
  run<4, 1>("lse v4 (a tree, bcd LUT) ILP1", 1, 1);
  run<4, 2>("lse v4 ILP2", 1, 1);
  run<4, 1>("lse v4 ILP1 16 warps", 16, 1);
  run<5, 1>("lse v5 (ab trees, cd LUT) ILP1", 1, 1);
  run<5, 2>("lse v5 ILP2", 1, 1);
  run<5, 1>("lse v5 ILP1 16 warps", 16, 1);
  run<3, 1>("lse v3 (parallel compares + LUT) ILP1", 1, 1);
  run<3, 3>("lse v3 ILP3", 1, 1);
  run<3, 1>("lse v3 ILP1 16 warps", 16, 1);
  return 0;
}
