// portable.h — lets the DP "phase" functions (scorers, numerics, per-diagonal phases) compile both as
// CUDA device code (the product) and as plain host C++ (tests/emu: a barrier-by-barrier emulator of the
// kernel used to check the parallel decomposition against the oracle without a GPU).  The host build
// defines the handful of CUDA intrinsics the phases use with identical IEEE-754 semantics; it must be
// compiled with -ffp-contract=off so that no a*b+c is fused (the device code uses __fmul_rn/__fadd_rn,
// which never contract).
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#include <cuda_runtime.h>
#define RNA_DEV __device__ __forceinline__
#define RNA_DEVM __device__ __forceinline__   // member functions
#define RNA_CONST_TABLE __device__ __constant__
#define RNA_PREFETCH_L2(p) asm volatile("prefetch.global.L2 [%0];" ::"l"(p))
// 4-byte asynchronous copy global -> shared (no register holds the value in flight), commit / wait of this thread's groups
#define RNA_CP_ASYNC4(sptr, gptr) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(sptr)), "l"(gptr) : "memory")
// (the same with a 32-bit shared-memory address computed by the caller)
#define RNA_CP_ASYNC4_S(saddr, gptr) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(saddr), "l"(gptr) : "memory")
#define RNA_CP_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define RNA_CP_WAIT(n) asm volatile("cp.async.wait_group %0;" ::"n"(n) : "memory")
// the two-loop term streams (st.global.cs / ld.global.cs measured 9 % SLOWER on the default bench: plain accesses)
#define RNA_ST_STREAM(p, v) (*(p) = (v))
#define RNA_LD_STREAM(p) (*(p))
// chains of the cooperative kernel; compiled as separate functions with -DRNA_COOP_NOINLINE (measured slower)
#ifdef RNA_COOP_NOINLINE
#define RNA_DEV_CALL __device__ __noinline__
#else
#define RNA_DEV_CALL __device__ __forceinline__
#endif
#else
#include <math.h>
#include <string.h>
#define RNA_DEV static inline
#define RNA_DEVM inline
#define RNA_CONST_TABLE static const
#define RNA_PREFETCH_L2(p) ((void)(p))
#define RNA_ST_STREAM(p, v) (*(p) = (v))
#define RNA_LD_STREAM(p) (*(p))
#define RNA_DEV_CALL static inline
#define __restrict__ __restrict
struct float4 { float x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { uint2 r; r.x = x; r.y = y; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __int_as_float(int x) { float f; memcpy(&f, &x, 4); return f; }
static inline int __float_as_int(float f) { int x; memcpy(&x, &f, 4); return x; }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline int __clz(unsigned x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) {
  const unsigned long long v = ((unsigned long long)hi << 32) | lo;
  return (unsigned)(v >> (sh & 31));
}
template <class T> static inline T min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T max(T a, T b) { return a > b ? a : b; }
#endif
