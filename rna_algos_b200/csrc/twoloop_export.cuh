// twoloop_export.cuh — FoldScores::twoloop_scores (src/mccaskill_algo.rs:13-22): the 4-D score memo that the reference's
// inside pass fills at :320 / :431 and returns to the caller.  The fold kernels keep these scores in their term streams
// (group-major, padded, per pass); this file lists them in the reference's own terms: for every pair (i,j) the inside
// pass visits, in its visiting order (span ascending, i ascending), the partners (k,l) that have a sums_close entry,
// k ascending and l descending (src/mccaskill_algo.rs:306-324, 412-435), each with get_2loop_score(seq,(i,j),(k,l)).
// Two kernels, one thread per cell: count, then (after an exclusive scan of the counts) fill.
#pragma once
#include "fold_kernel.cuh"
#include "fold_phases.cuh"

namespace rna {

struct TwoloopArgs {
  const uint8_t* sq;        // bases with 4 zero pads on both sides (points at base 0)
  const uint8_t* RR;        // RR[p] = s[p] * 4 + s[p + 1]
  const uint8_t* LL;        // LL[p] = s[p] * 4 + s[p - 1]
  int L;
  const void* tables;
  int allows_short;
  const float* close;       // RNA_SUMS_CLOSE plane of rna_fold_sums_batch: upper triangle incl. diagonal, -inf = no entry
  unsigned long long* offsets;   // per cell in visiting order: counts (count pass) / exclusive offsets (fill pass)
  RnaTwoloopScore* out;
  unsigned long long capacity;
};

struct TwoloopSeq {   // what the loop scorers read
  int L;
  const uint8_t* s;
  const uint8_t* RR;
  const uint8_t* LL;
};

__global__ void twoloop_codes_kernel(const uint8_t* bases, int L, uint8_t* sq_padded, uint8_t* RR, uint8_t* LLc) {
  const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
  for (int x = gt; x < L + 8; x += gn) {
    auto at = [&](int p) -> int { return (p >= 0 && p < L) ? bases[p] : 0; };
    const int p = x - 4;
    sq_padded[x] = (uint8_t)at(p);
    if (p >= 0 && p < L) { RR[p] = (uint8_t)(at(p) * 4 + at(p + 1)); LLc[p] = (uint8_t)(at(p) * 4 + at(p - 1)); }
  }
}

// visiting-order index of cell (d = j - i, i): diagonals ascending, i ascending
__host__ __device__ inline unsigned long long twoloop_cell(int L, int d, int i) {
  return (unsigned long long)d * (unsigned long long)L - (unsigned long long)d * (unsigned long long)(d - 1) / 2 + (unsigned long long)i;
}

template <bool CONTRA, bool FILL>
__global__ void twoloop_kernel(const TwoloopArgs a) {
  typedef typename Model2<CONTRA>::Dev Dev;
  typedef typename Model2<CONTRA>::View View;
  const Dev* dev = reinterpret_cast<const Dev*>(a.tables);
  View T;
  T.g = dev;
  T.sm = dev_small<CONTRA>(dev);
  const int L = a.L;
  const int MINSPAN = dev->min_span;
  int MAX2;
  if constexpr (CONTRA) MAX2 = dev->max_loop_len; else MAX2 = dev->max_2loop_len;
  TwoloopSeq v;
  v.L = L; v.s = a.sq; v.RR = a.RR; v.LL = a.LL;
  const unsigned long long ncell = (unsigned long long)L * (unsigned long long)(L + 1) / 2;
  const float NEG = RNA_NEG_INF;
  for (unsigned long long c = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += (unsigned long long)gridDim.x * blockDim.x) {
    // c -> (d, i): diagonal d holds L - d cells
    int d = (int)((2.0 * L + 1.0 - sqrt((2.0 * L + 1.0) * (2.0 * L + 1.0) - 8.0 * (double)c)) / 2.0);
    while (d > 0 && twoloop_cell(L, d, 0) > c) d--;
    while (twoloop_cell(L, d + 1, 0) <= c) d++;
    const int i = (int)(c - twoloop_cell(L, d, 0)), j = i + d;
    bool visits = canonical_pair(a.sq[i], a.sq[j]);
    if (CONTRA && a.allows_short) visits = visits && d >= 1; else visits = visits && (d + 1 >= MINSPAN);
    unsigned long long n = 0;
    const unsigned long long base = FILL ? a.offsets[c] : 0;
    if (visits) {
      const typename LoopOf<CONTRA, true>::type lp = make_loop<CONTRA, true>(v, T, i, j);
      for (int k = i + 1; k < j - 1; k++) {
        if (k - i - 1 > MAX2) break;
        for (int l = j - 1; l > k; l--) {
          if ((j - l - 1) + (k - i - 1) > MAX2) break;
          const float ckl = a.close[(unsigned long long)k * L - (unsigned long long)k * (k - 1) / 2 + (l - k)];
          if (!(ckl > NEG)) continue;
          if (FILL) {
            Term1 t1;
            t1.c = 0.f; t1.pv = 0.f; t1.q = 0; t1.a = k - i - 1; t1.b = j - l - 1; t1.code = a.RR[l] * 16 + a.LL[k];
            RnaTwoloopScore e;
            e.i = (uint16_t)i; e.j = (uint16_t)j; e.k = (uint16_t)k; e.l = (uint16_t)l;
            e.score = lp.score(lp.stage2(t1));
            if (base + n < a.capacity) a.out[base + n] = e;
          }
          n++;
        }
      }
    }
    if (!FILL) a.offsets[c] = n;
  }
}

}  // namespace rna
