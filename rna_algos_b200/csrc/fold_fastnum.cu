// fold_fastnum.cu — the FAST_F32 numeric mode's build of the batch kernel (include/rna_algos_b200.h
// RNA_NUMERIC_FAST_F32; sequences up to 1024 nt, one CTA each).
//
// Same kernel source as the reference-exact build (fold_kernel2.cuh / fold_phases.cuh: phases, warp roles, term
// streams, shared-memory / HBM-resident modes), compiled a second time with RNA_FASTNUM = 1: logsumexp and exp are
// exact log-space arithmetic on the MUFU pipe and the long chains accumulate in linear space (numerics.cuh,
// fold_phases.cuh ChainSum), so a chain step is one dependent FADD instead of a ~100-cycle polynomial logsumexp.
// Results agree with an exact-math evaluation of the reference's recurrences, not bit for bit with the reference
// (tolerances: DESIGN.md §2, tests/test_fast_mode.py).
//
// The second instantiation must not collide with the reference-exact one (same templates, different bodies), so this
// translation unit compiles the kernel headers into a namespace of its own; the launcher (rna_abi.cu) gets the kernel
// entry points through rna_fastnum_fold_kernel() and launches them with cudaLaunchKernel on the same FoldArgs image.
#define RNA_FASTNUM 1
#define rna rna_fastnum   // every `namespace rna` / `rna::` of the kernel headers below
#include "fold_kernel2.cuh"
#undef rna

extern "C" const void* rna_fastnum_fold_kernel(int contra, int hbm_mode) {
  using namespace rna_fastnum;
  if (contra) return hbm_mode ? (const void*)fold_kernel2<true, MODE_GLOBAL> : (const void*)fold_kernel2<true, MODE_SMEM>;
  return hbm_mode ? (const void*)fold_kernel2<false, MODE_GLOBAL> : (const void*)fold_kernel2<false, MODE_SMEM>;
}
extern "C" unsigned long rna_fastnum_sizeof_fold_args() { return (unsigned long)sizeof(rna_fastnum::FoldArgs); }
