// dev_tables.h — packed device images of the table blobs (SURVEY.md §8 "table packing").
// Built on the host from RnaTurnerTables / RnaContraTables / RnaAlignTables (rna_abi.cu) with plain
// IEEE f32 operations, so every precomputed entry has the bits the reference would compute in-line.
#pragma once
#include <stdint.h>

namespace rna {

// v2 kernels: the three interior-mismatch tables re-indexed by the per-position base codes
//   RR[p] = s[p]*4 + s[p+1],  LL[p] = s[p]*4 + s[p-1]:   tm2[X][RR*16 + LL] = TERMINAL_MISMATCH_X[x][y][x1][y1]
// (X = 0: 1xMANY, 1: 2x3, 2: INTERIOR), so one byte per partner position addresses them.
struct TurnerSmall2 {
  float stack[256];
  float tm2[3][256];
  float tm_hairpin[256];
  float tm_multi[256];
  float d5[64];
  float d3[64];
  float ninio[31 * 31];     // [a][b] = INTERIOR_SCORES_INIT[a+b] + max(NINIO_COEFF*|a-b|, NINIO_MAX)
  float bulge_init[32];
};

#define RNA_HAIRPIN_EXT_LEN 65536

struct DevTurner {
  int max_2loop_len;
  int min_span;
  int min_hairpin_len;
  int num_special;
  unsigned special_len_mask;        // bit n set <=> some special hairpin has slice length n (n < 32)
  float augu_pen;
  float init_mb_base;
  float coeff_num_branches;
  float bulge_init[31];
  float interior_init_ninio[31 * 31];   // [a][b] = INTERIOR_SCORES_INIT[a+b] + max(NINIO_COEFF*|a-b|, NINIO_MAX)
  unsigned special_key[128];        // 2 bits per base, base p at bits 2p (slice incl. closing pair)
  unsigned char special_len[128];
  float special_score[128];
  TurnerSmall2 small2;
  // device-global arrays
  const float* hairpin_init_ext;    // [RNA_HAIRPIN_EXT_LEN]: INIT[len] or the ln-extrapolation (src/utils.rs:178-184)
  const float* int11;               // [4^6]
  const float* int12;               // [4^7]
  const float* int22;               // [4^8]
};

// v2 kernels (fold_phases.cuh): per-lane gathers.  js / b1 / i11 / ptab are host-precomputed combinations;
// every entry is produced by the same f32 operations, in the same order, as the reference's scorer
// (src/utils.rs:456-520, 545-548), so using them is bit-identical to evaluating the scorer in-line.
struct ContraSmall2 {
  float js2[256];         // [RR*16 + LL] = helix_close[x][y] + terminal_mismatch[x][y][x1][y1]  (get_junction_score_single)
  float dl[64];
  float dr[64];
  float hc[16];
  float bp[16];
  float U[276 + 31 * 31];   // unified two-loop table, see below
};
// Unified two-loop table U: one gather per term.
//   [0,256)    stack_scores[i][j][k][l]
//   [256,260)  bulge_scores_0x1[x] + bulge_cum[0]                                              (bulge of length 1)
//   [260,276)  ((interior_1x1[x][y] + sym_cum[0]) + explicit[0][0]) + interior_cum[0]          (1x1 interior)
//   [276+a*31+b]  a+b>=2: bulge: 0 + bulge_cum[len-1]; interior: ((sym|asym) + explicit|0) + interior_cum[len-2]
#define RNA_CU_B1 256
#define RNA_CU_I11 260
#define RNA_CU_PTAB 276
#define RNA_CU_LEN (276 + 31 * 31)

struct DevContra {
  int max_loop_len;
  int min_span;
  int max_explicit;
  float mb_base, mb_bp, mb_unpair, ext_bp, ext_unpair;
  float mb_base_plus_bp;            // multibranch_score_base + multibranch_score_basepair (src/mccaskill_algo.rs:437-438)
  float hairpin_cum[31];
  float bulge_cum[30];
  float interior_cum[29];
  float sym_cum[15];
  float asym_cum[28];
  float explicit_[16];
  ContraSmall2 small2;
};

struct DevAlign {
  float m2m, m2i, iex, inm, ini;
  float insert[4];
  float match[16];
};

}  // namespace rna
