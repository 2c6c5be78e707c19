// table_pack.h — host-side packing of the C-ABI table blobs (include/rna_algos_b200.h) into the device
// images of dev_tables.h.  Pure host C++ (no CUDA calls): shared by rna_abi.cu and by the host emulator in
// tests/emu.  Every derived entry is computed with plain IEEE f32 operations in the reference's order
// (compile with -ffp-contract=off; `volatile` pins each rounding), so a precomputed entry has exactly the
// bits the reference would compute in-line.
#pragma once
#include <math.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/rna_algos_b200.h"
#include "dev_tables.h"

namespace rna {

static inline float f32_add(float a, float b) { volatile float r = a + b; return r; }
static inline float f32_mul(float a, float b) { volatile float r = a * b; return r; }

// Scores must be finite or -inf ("forbidden"): a NaN or +inf entry would leave the domain in which the kernels'
// branch-free logsumexp equals the reference's `if !x.is_finite() { return }` (src/utils.rs:581-586).
static inline bool scores_ok(const void* first, const void* end) {
  for (const float* p = static_cast<const float*>(first); p < static_cast<const float*>(end); p++)
    if (!(*p < INFINITY)) return false;   // NaN or +inf
  return true;
}

// Fills everything of DevTurner except the four device pointers; hp_ext receives the hairpin-initiation
// table for every loop length (table value, or the f32 ln-extrapolation of src/utils.rs:178-184).
static inline int pack_turner(const RnaTurnerTables* t, DevTurner* d, std::vector<float>* hp_ext, std::string* err) {
  if (t->max_2loop_len < 0 || t->max_2loop_len > 30 || t->min_span_hairpin_close < 2 ||
      t->min_hairpin_len < 0 || t->min_hairpin_len > 30 || t->max_hairpin_len_extrapolation > 30 ||
      t->max_hairpin_len_extrapolation < t->min_hairpin_len || t->min_hairpin_len_extrapolation < 2 ||
      t->min_hairpin_len_extrapolation > 31 || t->num_special_hairpins < 0 ||
      t->num_special_hairpins > RNA_MAX_SPECIAL_HAIRPINS) {
    *err = "Turner blob: caps out of range";
    return RNA_ERR_BAD_TABLES;
  }
  if (!scores_ok(&t->coeff_hairpin_len_extrapolation, &t->hairpin_scores_special[0])) {
    *err = "Turner blob: a score is NaN or +inf";
    return RNA_ERR_BAD_TABLES;
  }
  memset(d, 0, sizeof *d);
  d->max_2loop_len = t->max_2loop_len;
  d->min_span = t->min_span_hairpin_close;
  d->min_hairpin_len = t->min_hairpin_len;
  d->num_special = t->num_special_hairpins;
  d->augu_pen = t->helix_augu_end_penalty;
  d->init_mb_base = t->init_multibranch_base;
  d->coeff_num_branches = t->coeff_num_branches;
  memcpy(d->bulge_init, t->bulge_scores_init, sizeof d->bulge_init);
  for (int a = 0; a < 31; a++)
    for (int b = 0; b < 31; b++) {
      float v = 0.f;
      if (a + b <= 30) {
        const int diff = a > b ? a - b : b - a;
        const float nin = f32_mul(t->ninio_coeff, (float)diff);   // NINIO_COEFF * diff as Score
        const float nn = fmaxf(nin, t->ninio_max);                 // .max(NINIO_MAX), src/utils.rs:307
        v = f32_add(t->interior_scores_init[a + b], nn);
      }
      d->interior_init_ninio[a * 31 + b] = v;
    }
  for (int x = 0; x < t->num_special_hairpins; x++) {
    const RnaSpecialHairpin& e = t->hairpin_scores_special[x];
    if (e.len < 2 || e.len > RNA_MAX_SPECIAL_HAIRPIN_LEN) { *err = "Turner blob: special hairpin length"; return RNA_ERR_BAD_TABLES; }
    if (!(e.score < INFINITY)) { *err = "Turner blob: a special hairpin score is NaN or +inf"; return RNA_ERR_BAD_TABLES; }
    unsigned key = 0;
    for (int p = 0; p < e.len; p++) {
      if (e.seq[p] > 3) { *err = "Turner blob: special hairpin base"; return RNA_ERR_BAD_TABLES; }
      key |= (unsigned)e.seq[p] << (2 * p);
    }
    d->special_key[x] = key;
    d->special_len[x] = e.len;
    d->special_score[x] = e.score;
    d->special_len_mask |= 1u << e.len;
  }
  memcpy(d->small2.stack, t->stack_scores, 1024);
  memcpy(d->small2.tm_hairpin, t->terminal_mismatch_scores_hairpin, 1024);
  memcpy(d->small2.tm_multi, t->terminal_mismatch_scores_multibranch, 1024);
  memcpy(d->small2.d5, t->dangling_scores_5prime, 256);
  memcpy(d->small2.d3, t->dangling_scores_3prime, 256);
  for (int x = 0; x < 4; x++)
    for (int y = 0; y < 4; y++)
      for (int x1 = 0; x1 < 4; x1++)
        for (int y1 = 0; y1 < 4; y1++) {
          const int code = (x * 4 + x1) * 16 + (y * 4 + y1);
          d->small2.tm2[0][code] = t->terminal_mismatch_scores_1xmany[x][y][x1][y1];
          d->small2.tm2[1][code] = t->terminal_mismatch_scores_2x3[x][y][x1][y1];
          d->small2.tm2[2][code] = t->terminal_mismatch_scores_interior[x][y][x1][y1];
        }
  memcpy(d->small2.ninio, d->interior_init_ninio, sizeof d->small2.ninio);
  memcpy(d->small2.bulge_init, t->bulge_scores_init, sizeof(float) * 31);
  hp_ext->assign(RNA_HAIRPIN_EXT_LEN, 0.f);
  const int mex = t->min_hairpin_len_extrapolation - 1;
  for (int len = 0; len < RNA_HAIRPIN_EXT_LEN; len++) {
    if (len <= t->max_hairpin_len_extrapolation) {
      (*hp_ext)[len] = t->hairpin_scores_init[len];
    } else {
      volatile float ratio = (float)len / (float)mex;
      volatile float lg = logf(ratio);
      const float pr = f32_mul(t->coeff_hairpin_len_extrapolation, lg);
      (*hp_ext)[len] = f32_add(t->hairpin_scores_init[mex], pr);
    }
  }
  return RNA_OK;
}

static inline int pack_contra(const RnaContraTables* t, DevContra* d, std::string* err) {
  if (t->max_loop_len != RNA_CONTRA_MAX_LOOP_LEN || t->min_span_hairpin_close < 2 ||
      t->max_interior_explicit < 0 || t->max_interior_explicit > RNA_CONTRA_MAX_INTERIOR_EXPLICIT) {
    *err = "CONTRAfold blob: caps out of range";
    return RNA_ERR_BAD_TABLES;
  }
  if (!scores_ok(&t->hairpin_scores_len[0], t + 1)) {
    *err = "CONTRAfold blob: a score is NaN or +inf";
    return RNA_ERR_BAD_TABLES;
  }
  memset(d, 0, sizeof *d);
  d->max_loop_len = t->max_loop_len;
  d->min_span = t->min_span_hairpin_close;
  d->max_explicit = t->max_interior_explicit;
  d->mb_base = t->multibranch_score_base;
  d->mb_bp = t->multibranch_score_basepair;
  d->mb_unpair = t->multibranch_score_unpair;
  d->ext_bp = t->external_score_basepair;
  d->ext_unpair = t->external_score_unpair;
  d->mb_base_plus_bp = f32_add(t->multibranch_score_base, t->multibranch_score_basepair);
  memcpy(d->hairpin_cum, t->hairpin_scores_len_cumulative, sizeof d->hairpin_cum);
  memcpy(d->bulge_cum, t->bulge_scores_len_cumulative, sizeof d->bulge_cum);
  memcpy(d->interior_cum, t->interior_scores_len_cumulative, sizeof d->interior_cum);
  memcpy(d->sym_cum, t->interior_scores_symmetric_cumulative, sizeof d->sym_cum);
  memcpy(d->asym_cum, t->interior_scores_asymmetric_cumulative, sizeof d->asym_cum);
  memcpy(d->explicit_, t->interior_scores_explicit, sizeof d->explicit_);
  // ---- v2 combinations --------------------------------------------------------------------------
  ContraSmall2& s2 = d->small2;
  memcpy(s2.dl, t->dangling_scores_left, 256);
  memcpy(s2.dr, t->dangling_scores_right, 256);
  memcpy(s2.hc, t->helix_close_scores, 64);
  memcpy(s2.bp, t->basepair_scores, 64);
  for (int x = 0; x < 4; x++)
    for (int y = 0; y < 4; y++)
      for (int x1 = 0; x1 < 4; x1++)
        for (int y1 = 0; y1 < 4; y1++)   // get_junction_score_single, src/utils.rs:545-548
          s2.js2[(x * 4 + x1) * 16 + (y * 4 + y1)] = f32_add(t->helix_close_scores[x][y], t->terminal_mismatch_scores[x][y][x1][y1]);
  const int me = t->max_interior_explicit;
  memcpy(s2.U, t->stack_scores, 1024);
  for (int x = 0; x < 4; x++)       // src/utils.rs:464-474: score(0x1) + bulge_cum[len-1], len == 1
    s2.U[RNA_CU_B1 + x] = f32_add(t->bulge_scores_0x1[x], t->bulge_scores_len_cumulative[0]);
  for (int x = 0; x < 4; x++)
    for (int y = 0; y < 4; y++) {   // src/utils.rs:495-514 with loop_len_pair == (1,1)
      float v = f32_add(t->interior_scores_1x1[x][y], t->interior_scores_symmetric_cumulative[0]);
      v = f32_add(v, me >= 1 ? t->interior_scores_explicit[0][0] : 0.f);
      v = f32_add(v, t->interior_scores_len_cumulative[0]);
      s2.U[RNA_CU_I11 + x * 4 + y] = v;
    }
  for (int a = 0; a < 31; a++)
    for (int b = 0; b < 31; b++) {
      float v = 0.f;
      const int len = a + b;
      if (len >= 2 && len <= 30) {
        if (a == 0 || b == 0) {
          v = f32_add(0.f, t->bulge_scores_len_cumulative[len - 1]);             // src/utils.rs:464-476, len >= 2
        } else if (!(a == 1 && b == 1)) {
          if (a == b) v = f32_add(0.f, t->interior_scores_symmetric_cumulative[a - 1]);
          else v = t->interior_scores_asymmetric_cumulative[(a > b ? a - b : b - a) - 1];
          v = f32_add(v, (a <= me && b <= me) ? t->interior_scores_explicit[a - 1][b - 1] : 0.f);
          v = f32_add(v, t->interior_scores_len_cumulative[len - 2]);
        }
      }
      s2.U[RNA_CU_PTAB + a * 31 + b] = v;
    }
  return RNA_OK;
}

static inline int pack_align(const RnaAlignTables* t, DevAlign* d, std::string* err) {
  if (!scores_ok(t, t + 1)) {
    *err = "align blob: a score is NaN or +inf";
    return RNA_ERR_BAD_TABLES;
  }
  d->m2m = t->match2match_score;
  d->m2i = t->match2insert_score;
  d->iex = t->insert_extend_score;
  d->inm = t->init_match_score;
  d->ini = t->init_insert_score;
  memcpy(d->insert, t->insert_scores, 16);
  memcpy(d->match, t->match_scores, 64);
  return RNA_OK;
}

}  // namespace rna
