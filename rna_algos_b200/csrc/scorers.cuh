// scorers.cuh — Turner 2004 / CONTRAfold v2.02 loop scorers (reference src/utils.rs:166-556) over the
// packed device tables.  Portable (see portable.h): compiled by nvcc for the kernels and by g++ for the
// host emulator in tests/emu.
#pragma once
#include "dev_tables.h"
#include "numerics.cuh"

namespace rna {

RNA_DEV int idx4(int a, int b, int c, int d) { return ((a * 4 + b) * 4 + c) * 4 + d; }
RNA_DEV int idx3(int a, int b, int c) { return (a * 4 + b) * 4 + c; }

// ---------------------------------------------------------------------------------------------------
// Turner 2004 scorers (src/utils.rs:166-411).  `s` = sequence bytes (shared memory), T = a view {g: device tables
// (uniform reads), sm: their shared-memory part (per-lane gathers)}: TurnerView2 of fold_phases.cuh.
// ---------------------------------------------------------------------------------------------------
template <class TV>
RNA_DEV float t_pen(const TV& T, int x, int y) {
  return augu_pair(x, y) ? T.g->augu_pen : 0.f;
}

// get_hairpin_score, src/utils.rs:166-196 (special-loop scan :198-205 via packed keys)
template <class TV>
RNA_DEV float t_hairpin(const TV& T, const uint8_t* s, int i, int j) {
  const int span = j - i + 1;
  if (span < 32 && ((T.g->special_len_mask >> span) & 1u)) {
    unsigned key = 0;
    for (int p = 0; p < span; p++) key |= (unsigned)s[i + p] << (2 * p);
    const int n = T.g->num_special;
    for (int x = 0; x < n; x++) {
      if ((int)T.g->special_len[x] == span && T.g->special_key[x] == key) {
        // first match decides (src/utils.rs:198-205); a listed score of -inf falls through to the generic
        // hairpin (`if special_hairpin_score > NEG_INFINITY`, src/utils.rs:169)
        if (T.g->special_score[x] > RNA_NEG_INF) return T.g->special_score[x];
        break;
      }
    }
  }
  const int len = j - i - 1;
  const int si = s[i], sj = s[j];
  float hs;
  if (len == T.g->min_hairpin_len) {
    hs = __ldg(&T.g->hairpin_init_ext[len]);
  } else {
    hs = __fadd_rn(__ldg(&T.g->hairpin_init_ext[len]), T.sm->tm_hairpin[idx4(si, sj, s[i + 1], s[j - 1])]);
  }
  return __fadd_rn(hs, t_pen(T, si, sj));
}

// get_multibranch_close_score, src/utils.rs:368-382
template <class TV>
RNA_DEV float t_mbclose(const TV& T, const uint8_t* s, int i, int j) {
  const int si = s[i], sj = s[j];
  const float tm = T.sm->tm_multi[idx4(sj, si, s[j - 1], s[i + 1])];
  return __fadd_rn(__fadd_rn(T.g->init_mb_base, tm), t_pen(T, si, sj));
}

// get_accessible_score (uses_sentinel_bases = false), src/utils.rs:384-411
template <class TV>
RNA_DEV float t_acc(const TV& T, const uint8_t* s, int L, int i, int j) {
  const int si = s[i], sj = s[j];
  float sc;
  if (i > 0 && j < L - 1) sc = T.sm->tm_multi[idx4(si, sj, s[i - 1], s[j + 1])];
  else if (i > 0) sc = T.sm->d5[idx3(si, sj, s[i - 1])];
  else if (j < L - 1) sc = T.sm->d3[idx3(si, sj, s[j + 1])];
  else sc = 0.f;
  return __fadd_rn(sc, t_pen(T, si, sj));
}

// (The CONTRAfold scorers, src/utils.rs:413-556, and the Turner two-loop scorer work on host-precomputed combination
// tables addressed by per-position base codes: fold_phases.cuh.)

}  // namespace rna
