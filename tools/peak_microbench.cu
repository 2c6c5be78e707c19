// peak_microbench.cu — measured denominators of bench.py's roofline that MEASURED_PEAKS.json does not hold
// (BASELINE.md §3): the FP32 NON-FMA issue rate (the reference-exact path may not contract a*b+c, so
// an FMA counts as nothing here: the peak is thread-instructions/s of separate FMUL / FADD), the MUFU (ex2.approx)
// rate the FAST numeric mode leans on, and the shared-memory load bandwidth.  Built as a small shared library
// (tools/_build/libpeaks.so, C ABI) that bench.py loads and runs in the SAME process as the timed kernels, so the
// clocks are the ones the kernels see.  Measurement infrastructure: not linked by the product library.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -shared -Xcompiler -fPIC -o libpeaks.so peak_microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>

namespace {
constexpr int kIters = 4096;

// 8 independent chains per thread, alternating FMUL / FADD (never fused): 16 FP32 instructions per iteration
__global__ void __launch_bounds__(256) fp32_nofma_kernel(float* out, float a, float b) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
#pragma unroll 4
  for (int it = 0; it < kIters; it++) {
    x0 = __fmul_rn(x0, a); x1 = __fmul_rn(x1, a); x2 = __fmul_rn(x2, a); x3 = __fmul_rn(x3, a);
    x4 = __fmul_rn(x4, a); x5 = __fmul_rn(x5, a); x6 = __fmul_rn(x6, a); x7 = __fmul_rn(x7, a);
    x0 = __fadd_rn(x0, b); x1 = __fadd_rn(x1, b); x2 = __fadd_rn(x2, b); x3 = __fadd_rn(x3, b);
    x4 = __fadd_rn(x4, b); x5 = __fadd_rn(x5, b); x6 = __fadd_rn(x6, b); x7 = __fadd_rn(x7, b);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

__device__ __forceinline__ float ex2a(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// 8 independent ex2.approx chains per thread: 8 MUFU instructions per iteration (plus 8 FADDs that keep the values bounded)
__global__ void __launch_bounds__(256) mufu_kernel(float* out, float c) {
  float x0 = threadIdx.x * 1e-3f, x1 = x0 + .1f, x2 = x0 + .2f, x3 = x0 + .3f, x4 = x0 + .4f, x5 = x0 + .5f, x6 = x0 + .6f, x7 = x0 + .7f;
#pragma unroll 4
  for (int it = 0; it < kIters; it++) {
    x0 = ex2a(x0) + c; x1 = ex2a(x1) + c; x2 = ex2a(x2) + c; x3 = ex2a(x3) + c;
    x4 = ex2a(x4) + c; x5 = ex2a(x5) + c; x6 = ex2a(x6) + c; x7 = ex2a(x7) + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}

// conflict-free 16-byte shared-memory loads, 8 in flight per thread
__global__ void __launch_bounds__(256) smem_kernel(float* out) {
  __shared__ float4 buf[2048];   // 32 KB
  for (int x = threadIdx.x; x < 2048; x += blockDim.x) buf[x] = make_float4(x, 1.f, 2.f, 3.f);
  __syncthreads();
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  int p = threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < kIters / 8; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      const float4 v = buf[(p + 256 * u) & 2047];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    p = (p + 17) & 2047;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (acc.x + acc.y) + (acc.z + acc.w);
}

template <class F>
double timed_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  return best;
}
}  // namespace

// out[0] = FP32 non-FMA thread-instructions/s, out[1] = MUFU (ex2) thread-instructions/s, out[2] = shared-memory load B/s,
// out[3] = SM count.  Returns 0, or a CUDA error code.
extern "C" int peaks_measure(int device, double* out) {
  if (cudaSetDevice(device) != cudaSuccess) return 1;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return 2;
  const int grid = prop.multiProcessorCount * 8, nt = 256;   // 2048 threads per SM
  float* d = nullptr;
  if (cudaMalloc(&d, sizeof(float) * grid * nt) != cudaSuccess) return 3;
  const double n_threads = (double)grid * nt;
  double ms = timed_ms([&] { fp32_nofma_kernel<<<grid, nt>>>(d, 1.0000001f, 1e-7f); });
  out[0] = n_threads * kIters * 16.0 / (ms * 1e-3);
  ms = timed_ms([&] { mufu_kernel<<<grid, nt>>>(d, -0.999f); });
  out[1] = n_threads * kIters * 8.0 / (ms * 1e-3);
  ms = timed_ms([&] { smem_kernel<<<grid, nt>>>(d); });
  out[2] = n_threads * kIters * 16.0 / (ms * 1e-3);
  out[3] = prop.multiProcessorCount;
  cudaFree(d);
  return cudaGetLastError() == cudaSuccess ? 0 : 4;
}
