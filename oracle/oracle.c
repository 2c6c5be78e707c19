/*
 * oracle.c — CPU restatement of the rna-algos hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle (and the `cpu_baseline` / `--impl reference` arm of bench.py).  It is
 * never linked, imported or executed by the product (rna_algos_b200/); only tests/,
 * __graft_entry__.smoke() and bench.py's CPU legs may use it.
 *
 * PARITY UNPINNED: the reference (heartsh/rna-algos 0.1.37) is Rust and cannot be built here (no
 * rustc/cargo, its table crate `rna-ss-params = "0.1"` is not vendored and there is no Cargo.lock),
 * and its own tests hold no golden vectors (tests/tests.rs:33,38,74 assert only p in [-0.001, 1.001)).
 * What pins this file instead: (1) it follows the reference line by line — every function cites the
 * file:line it restates, in the reference's loop order, fold order and f32 rounding (compile with
 * -ffp-contract=off); (2) an f64 exact-math flavour of the same recurrences (-DORC_EXACT) is checked
 * against brute-force enumeration of all secondary structures (tests/test_oracle_bruteforce.py);
 * (3) the reference's own range assertion; (4) the CONTRAlign constants are the in-tree ones.
 *
 * Dense L x L arrays with -inf for "key absent" replace the reference's hash maps: a key is inserted
 * iff its value > -inf (src/mccaskill_algo.rs:332-342, 456-466, 602-604), so presence == (v > -inf).
 * The two-loop score memo (fold_scores.twoloop_scores, :320,:589) is a pure function and is recomputed.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rna_algos_b200.h"

#ifdef ORC_EXACT
typedef double real;
#define R_LOG log
#define R_EXP exp
#else
typedef float real;
#define R_LOG logf
#define R_EXP expf
#endif

#define NEG_INF ((real)(-INFINITY))

/* Number of logsumexp calls with a finite operand ("LSE-terms"). */
static __thread uint64_t g_lse_terms;

/* ---------------------------------------------------------------------------------------------
 * Numerics: src/utils.rs:579-655
 * ------------------------------------------------------------------------------------------- */

/* src/utils.rs:603-627.  The reference's comparison tree picks one of eight cubic segments; here the segment index
 * is counted from the same seven breakpoints (`!(x < b)`, so a NaN lands in the last segment like in the tree) and
 * the Horner form is evaluated with that segment's coefficients: same operations, same order, no data-dependent
 * branch (the tree mispredicts on random operands and dominated the run time of long sequences). */
#ifndef ORC_EXACT
static const float kLnExp1pBreaks[7] = {0.66153675f, 1.6320158f, 2.4912589f, 3.37925f, 4.426169f, 5.789071f, 7.8162727f};
static const float kLnExp1pCoef[8][4] = {
    {-0.0065591595f, 0.12764427f, 0.49965546f, 0.6931542f},   {-0.015515756f, 0.14467756f, 0.48829398f, 0.6958093f},
    {-0.012890925f, 0.13010283f, 0.51503986f, 0.6795586f},    {-0.0072142647f, 0.087754086f, 0.6208708f, 0.5909676f},
    {-0.0031455354f, 0.046722945f, 0.7592532f, 0.43487945f},  {-0.0010110698f, 0.018594341f, 0.88317305f, 0.25236955f},
    {-0.000196278f, 0.0046084408f, 0.9634432f, 0.09831489f},  {-0.0000113994f, 0.0003734731f, 0.9959107f, 0.0149855051f}};
#endif
static inline real ln_exp_1p(real x) {
#ifdef ORC_EXACT
  return log1p(exp(x));
#else
  int seg = 0;
  for (int b = 0; b < 7; b++) seg += !(x < kLnExp1pBreaks[b]);
  const float *c = kLnExp1pCoef[seg];
  return ((c[0] * x + c[1]) * x + c[2]) * x + c[3];
#endif
}

/* src/utils.rs:580-596 */
static inline void lse(real *sum, real x) {
  if (!isfinite(x)) return;
  g_lse_terms++;
  if (!isfinite(*sum)) {
    *sum = x;
  } else {
    real y = *sum < x ? *sum : x;            /* sum.min(x) */
    real z = (*sum > x ? *sum : x) - y;      /* sum.max(x) - y */
#ifdef ORC_EXACT
    *sum = y + (z + log1p(exp(-z)));
#else
    *sum = y + (z >= 11.862479f ? z : ln_exp_1p(z));
#endif
  }
}

/* src/utils.rs:631-655 */
static inline real approx_expf(real x) {
#ifdef ORC_EXACT
  return exp(x);
#else
  if (x < -2.4915035f) {
    if (x < -5.8622823f) {
      if (x < -9.91152f) {
        return 0.f;
      } else {
        return ((0.0000803850f * x + 0.002162743f) * x + 0.019470856f) * x + 0.058808003f;
      }
    } else if (x < -3.839663f) {
      return ((0.0013889414f * x + 0.024467647f) * x + 0.14712906f) * x + 0.30427578f;
    } else {
      return ((0.0072335607f * x + 0.09060027f) * x + 0.39831114f) * x + 0.62459594f;
    }
  } else if (x < -0.6725053f) {
    if (x < -1.4805375f) {
      return ((0.023241036f * x + 0.2085646f) * x + 0.6906368f) * x + 0.86823225f;
    } else {
      return ((0.057378277f * x + 0.35802585f) * x + 0.9121133f) * x + 0.9793092f;
    }
  } else if (x < 0.f) {
    return ((0.119917594f * x + 0.48156682f) * x + 0.9975992f) * x + 0.9999505f;
  } else {
    return R_EXP(x);
  }
#endif
}

/* Exposed for tests of the polynomial kernels. */
float orc_ln_exp_1p(float x) { return (float)ln_exp_1p((real)x); }
float orc_expf(float x) { return (float)approx_expf((real)x); }
float orc_logsumexp(float sum, float x) {
  real s = (real)sum;
  lse(&s, (real)x);
  return (float)s;
}

/* ---------------------------------------------------------------------------------------------
 * Base-pair predicates: src/utils.rs:162-164, 558-560
 * ------------------------------------------------------------------------------------------- */
static inline int canonical(int x, int y) {
  return (x == 0 && y == 3) || (x == 1 && y == 2) || (x == 2 && y == 1) || (x == 2 && y == 3) ||
         (x == 3 && y == 0) || (x == 3 && y == 2);
}
static inline int matches_augu(int x, int y) {
  return (x == 0 && y == 3) || (x == 3 && y == 0) || (x == 2 && y == 3) || (x == 3 && y == 2);
}

/* ---------------------------------------------------------------------------------------------
 * Turner 2004 loop scorers: src/utils.rs:166-411
 * ------------------------------------------------------------------------------------------- */
typedef const RnaTurnerTables TT;

/* src/utils.rs:198-205 — first list entry whose slice equals the whole hairpin (closing pair included). */
static real special_hairpin_score(const uint8_t *hp, int n, TT *t) {
  for (int x = 0; x < t->num_special_hairpins; x++) {
    const RnaSpecialHairpin *e = &t->hairpin_scores_special[x];
    if ((int)e->len == n && memcmp(e->seq, hp, (size_t)n) == 0) return (real)e->score;
  }
  return NEG_INF;
}

/* src/utils.rs:166-196 */
static real hairpin_score(const uint8_t *s, int i, int j, TT *t) {
  real sp = special_hairpin_score(s + i, j - i + 1, t);
  if (sp > NEG_INF) return sp;
  int len = j - i - 1;
  real hs;
  if (len == t->min_hairpin_len) {
    hs = (real)t->hairpin_scores_init[len];
  } else {
    real init;
    if (len <= t->max_hairpin_len_extrapolation) {
      init = (real)t->hairpin_scores_init[len];
    } else {
      init = (real)t->hairpin_scores_init[t->min_hairpin_len_extrapolation - 1] +
             (real)t->coeff_hairpin_len_extrapolation *
                 R_LOG((real)len / (real)(t->min_hairpin_len_extrapolation - 1));
    }
    hs = init + (real)t->terminal_mismatch_scores_hairpin[s[i]][s[j]][s[i + 1]][s[j - 1]];
  }
  return hs + (matches_augu(s[i], s[j]) ? (real)t->helix_augu_end_penalty : (real)0.);
}

/* src/utils.rs:224-232 */
static real stack_score(const uint8_t *s, int i, int j, int k, int l, TT *t) {
  return (real)t->stack_scores[s[i]][s[j]][s[k]][s[l]];
}

/* src/utils.rs:234-258 */
static real bulge_score(const uint8_t *s, int i, int j, int k, int l, TT *t) {
  int len = k - i + j - l - 2;
  if (len == 1) {
    return (real)t->bulge_scores_init[len] + stack_score(s, i, j, k, l, t);
  }
  return (real)t->bulge_scores_init[len] +
         (matches_augu(s[i], s[j]) ? (real)t->helix_augu_end_penalty : (real)0.) +
         (matches_augu(s[k], s[l]) ? (real)t->helix_augu_end_penalty : (real)0.);
}

/* src/utils.rs:331-366 */
static real interior_mismatch_score(const uint8_t *s, int i, int j, int k, int l, int a, int b, TT *t) {
  int c0 = s[i], c1 = s[j];   /* basepair_close */
  int a0 = s[l], a1 = s[k];   /* basepair_accessible = (seq[acc.1], seq[acc.0]) */
  int m00 = s[i + 1], m01 = s[j - 1];
  int m10 = s[l + 1], m11 = s[k - 1];
  if (a == 1 || b == 1) {
    return (real)t->terminal_mismatch_scores_1xmany[c0][c1][m00][m01] +
           (real)t->terminal_mismatch_scores_1xmany[a0][a1][m10][m11];
  } else if ((a == 2 && b == 3) || (a == 3 && b == 2)) {
    return (real)t->terminal_mismatch_scores_2x3[c0][c1][m00][m01] +
           (real)t->terminal_mismatch_scores_2x3[a0][a1][m10][m11];
  }
  return (real)t->terminal_mismatch_scores_interior[c0][c1][m00][m01] +
         (real)t->terminal_mismatch_scores_interior[a0][a1][m10][m11];
}

/* src/utils.rs:260-321 */
static real interior_score(const uint8_t *s, int i, int j, int k, int l, TT *t) {
  int a = k - i - 1, b = j - l - 1;
  if (a == 1 && b == 1) {
    return (real)t->interior_scores_1x1[s[i]][s[j]][s[i + 1]][s[j - 1]][s[k]][s[l]];
  } else if (a == 1 && b == 2) {
    return (real)t->interior_scores_1x2[s[i]][s[j]][s[i + 1]][s[j - 1]][s[j - 2]][s[k]][s[l]];
  } else if (a == 2 && b == 1) {
    /* accessible pair inverted first, close pair inverted last (src/utils.rs:286-296) */
    return (real)t->interior_scores_1x2[s[l]][s[k]][s[j - 1]][s[i + 2]][s[i + 1]][s[j]][s[i]];
  } else if (a == 2 && b == 2) {
    return (real)t->interior_scores_2x2[s[i]][s[j]][s[i + 1]][s[j - 1]][s[i + 2]][s[j - 2]][s[k]][s[l]];
  }
  int diff = a > b ? a - b : b - a;
  real ninio = (real)t->ninio_coeff * (real)diff;
  if (!(ninio > (real)t->ninio_max)) ninio = (real)t->ninio_max;   /* .max(NINIO_MAX) */
  return (real)t->interior_scores_init[a + b] + ninio +
         interior_mismatch_score(s, i, j, k, l, a, b, t) +
         (matches_augu(s[i], s[j]) ? (real)t->helix_augu_end_penalty : (real)0.) +
         (matches_augu(s[k], s[l]) ? (real)t->helix_augu_end_penalty : (real)0.);
}

/* src/utils.rs:207-222 */
static real twoloop_score(const uint8_t *s, int i, int j, int k, int l, TT *t) {
  if (i + 1 == k && j - 1 == l) return stack_score(s, i, j, k, l, t);
  if (i + 1 == k || j - 1 == l) return bulge_score(s, i, j, k, l, t);
  return interior_score(s, i, j, k, l, t);
}

/* src/utils.rs:368-382 */
static real multibranch_close_score(const uint8_t *s, int i, int j, TT *t) {
  real tm = (real)t->terminal_mismatch_scores_multibranch[s[j]][s[i]][s[j - 1]][s[i + 1]];
  return (real)t->init_multibranch_base + tm +
         (matches_augu(s[i], s[j]) ? (real)t->helix_augu_end_penalty : (real)0.);
}

/* src/utils.rs:384-411 with uses_sentinel_bases = false */
static real accessible_score(const uint8_t *s, int L, int i, int j, TT *t) {
  int end5 = 0, end3 = L - 1;
  real sc;
  if (i > end5 && j < end3) {
    sc = (real)t->terminal_mismatch_scores_multibranch[s[i]][s[j]][s[i - 1]][s[j + 1]];
  } else if (i > end5) {
    sc = (real)t->dangling_scores_5prime[s[i]][s[j]][s[i - 1]];
  } else if (j < end3) {
    sc = (real)t->dangling_scores_3prime[s[i]][s[j]][s[j + 1]];
  } else {
    sc = (real)0.;
  }
  return sc + (matches_augu(s[i], s[j]) ? (real)t->helix_augu_end_penalty : (real)0.);
}

/* ---------------------------------------------------------------------------------------------
 * CONTRAfold v2.02 loop scorers: src/utils.rs:413-556
 * ------------------------------------------------------------------------------------------- */
typedef const RnaContraTables CT;

/* src/utils.rs:545-556 */
static real c_junction_single(const uint8_t *s, int p0, int p1, CT *t) {
  return (real)t->helix_close_scores[s[p0]][s[p1]] +
         (real)t->terminal_mismatch_scores[s[p0]][s[p1]][s[p0 + 1]][s[p1 - 1]];
}

/* src/utils.rs:522-543 with uses_sentinel_bases = false */
static real c_junction(const uint8_t *s, int L, int p0, int p1, CT *t) {
  int end5 = 0, end3 = L - 1;
  return (real)t->helix_close_scores[s[p0]][s[p1]] +
         (p0 < end3 ? (real)t->dangling_scores_left[s[p0]][s[p1]][s[p0 + 1]] : (real)0.) +
         (p1 > end5 ? (real)t->dangling_scores_right[s[p0]][s[p1]][s[p1 - 1]] : (real)0.);
}

/* src/utils.rs:413-421 */
static real c_hairpin_score(const uint8_t *s, int i, int j, CT *t) {
  int len = j - i - 1;
  int idx = len < t->max_loop_len ? len : t->max_loop_len;
  return (real)t->hairpin_scores_len_cumulative[idx] + c_junction_single(s, i, j, t);
}

/* src/utils.rs:456-481 */
static real c_bulge_score(const uint8_t *s, int i, int j, int k, int l, CT *t) {
  int len = k - i + j - l - 2;
  real sc = len == 1 ? (real)t->bulge_scores_0x1[(k - i - 1 == 1) ? s[i + 1] : s[j - 1]] : (real)0.;
  return sc + (real)t->bulge_scores_len_cumulative[len - 1] + c_junction_single(s, i, j, t) +
         c_junction_single(s, l, k, t);
}

/* src/utils.rs:483-520 */
static real c_interior_score(const uint8_t *s, int i, int j, int k, int l, CT *t) {
  int a = k - i - 1, b = j - l - 1;
  int len = a + b;
  real sc;
  if (a == b) {
    real s11 = len == 2 ? (real)t->interior_scores_1x1[s[i + 1]][s[j - 1]] : (real)0.;
    sc = s11 + (real)t->interior_scores_symmetric_cumulative[a - 1];
  } else {
    int diff = a > b ? a - b : b - a;
    sc = (real)t->interior_scores_asymmetric_cumulative[diff - 1];
  }
  real ex = (a <= t->max_interior_explicit && b <= t->max_interior_explicit)
                ? (real)t->interior_scores_explicit[a - 1][b - 1]
                : (real)0.;
  return sc + ex + (real)t->interior_scores_len_cumulative[len - 2] + c_junction_single(s, i, j, t) +
         c_junction_single(s, l, k, t);
}

/* src/utils.rs:423-442 */
static real c_twoloop_score(const uint8_t *s, int i, int j, int k, int l, CT *t) {
  real sc;
  if (i + 1 == k && j - 1 == l) {
    sc = (real)t->stack_scores[s[i]][s[j]][s[k]][s[l]];   /* src/utils.rs:444-454 */
  } else if (i + 1 == k || j - 1 == l) {
    sc = c_bulge_score(s, i, j, k, l, t);
  } else {
    sc = c_interior_score(s, i, j, k, l, t);
  }
  return sc + (real)t->basepair_scores[s[k]][s[l]];
}

/* ---------------------------------------------------------------------------------------------
 * Inside / outside state: FoldSums, src/mccaskill_algo.rs:3-11, 213-226
 * ------------------------------------------------------------------------------------------- */

/* ---------------------------------------------------------------------------------------------
 * Optional parallelism INSIDE one sequence (long sequences: the 2048 / 4096 nt parity checks).  The cells of one span
 * (anti-diagonal) are independent of one another — each reads only cells of other spans and writes its own — so
 * span_for() spreads them over a small persistent pthread team.  Every cell's folds run exactly as in the serial
 * loop (same operands, same order), so all values are bit-identical to the serial run.  Off by default: the
 * reference is serial per sequence (src/mccaskill_algo.rs:282-723) and so is the CPU baseline;
 * orc_set_inner_threads(n) turns it on for single-sequence calls of the calling process.
 * ------------------------------------------------------------------------------------------- */
typedef void (*span_cell_fn)(void *ctx, int i);
static int g_inner_threads = 1;
static struct {
  pthread_mutex_t mu;
  pthread_cond_t go, done;
  pthread_t *th;
  int n_workers;          /* helper threads (the caller works too) */
  uint64_t epoch;         /* bumped per parallel span */
  int running;            /* helpers still inside the current span */
  span_cell_fn fn;
  void *ctx;
  int ncells;
  volatile int next;      /* next chunk start */
  uint64_t terms;         /* LSE-terms counted by the helpers in the current span */
  int quit;
} g_team = {PTHREAD_MUTEX_INITIALIZER, PTHREAD_COND_INITIALIZER, PTHREAD_COND_INITIALIZER, NULL, 0, 0, 0, NULL, NULL, 0, 0, 0, 0};
#define SPAN_CHUNK 4

static void span_drain(void) {
  for (;;) {
    int b = __atomic_fetch_add(&g_team.next, SPAN_CHUNK, __ATOMIC_RELAXED);
    if (b >= g_team.ncells) break;
    int e = b + SPAN_CHUNK < g_team.ncells ? b + SPAN_CHUNK : g_team.ncells;
    for (int i = b; i < e; i++) g_team.fn(g_team.ctx, i);
  }
}
static void *span_helper(void *arg) {
  (void)arg;
  uint64_t seen = 0;
  pthread_mutex_lock(&g_team.mu);
  for (;;) {
    while (g_team.epoch == seen && !g_team.quit) pthread_cond_wait(&g_team.go, &g_team.mu);
    if (g_team.quit) break;
    seen = g_team.epoch;
    pthread_mutex_unlock(&g_team.mu);
    const uint64_t t0 = g_lse_terms;
    span_drain();
    const uint64_t dt = g_lse_terms - t0;
    pthread_mutex_lock(&g_team.mu);
    g_team.terms += dt;
    if (--g_team.running == 0) pthread_cond_signal(&g_team.done);
  }
  pthread_mutex_unlock(&g_team.mu);
  return NULL;
}
void orc_set_inner_threads(int n) {
  if (n < 1) n = 1;
  pthread_mutex_lock(&g_team.mu);
  if (g_team.th) {   /* stop the old team */
    g_team.quit = 1;
    pthread_cond_broadcast(&g_team.go);
    pthread_mutex_unlock(&g_team.mu);
    for (int x = 0; x < g_team.n_workers; x++) pthread_join(g_team.th[x], NULL);
    pthread_mutex_lock(&g_team.mu);
    free(g_team.th);
    g_team.th = NULL; g_team.n_workers = 0; g_team.quit = 0;
  }
  g_inner_threads = n;
  if (n > 1) {
    g_team.n_workers = n - 1;
    g_team.th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)(n - 1));
    for (int x = 0; x < n - 1; x++) pthread_create(&g_team.th[x], NULL, span_helper, NULL);
  }
  pthread_mutex_unlock(&g_team.mu);
}
/* fn(ctx, i) for i = 0 .. ncells-1: serial, or over the team (the caller takes part and returns when all are done) */
static void span_for(int ncells, span_cell_fn fn, void *ctx) {
  if (g_inner_threads <= 1 || ncells < 4 * SPAN_CHUNK) {
    for (int i = 0; i < ncells; i++) fn(ctx, i);
    return;
  }
  pthread_mutex_lock(&g_team.mu);
  g_team.fn = fn; g_team.ctx = ctx; g_team.ncells = ncells; g_team.next = 0; g_team.terms = 0;
  g_team.running = g_team.n_workers;
  g_team.epoch++;
  pthread_cond_broadcast(&g_team.go);
  pthread_mutex_unlock(&g_team.mu);
  span_drain();
  pthread_mutex_lock(&g_team.mu);
  while (g_team.running > 0) pthread_cond_wait(&g_team.done, &g_team.mu);
  g_lse_terms += g_team.terms;
  pthread_mutex_unlock(&g_team.mu);
}

/* what a cell function sees */
typedef struct {
  const uint8_t *s;
  int L, span, allows_short;
  const RnaTurnerTables *tt;
  const RnaContraTables *ct;
  struct FoldSums_ *f;
  float gamma_unused;
  void *bp, *pm, *pm2;     /* real* */
  double global_sum;       /* holds a `real` exactly */
} SpanCtx;

typedef struct FoldSums_ {
  int L;
  real *ext;    /* sums_external                          init 0.0 (incl. lower triangle) */
  real *rbe;    /* sums_rightmost_basepairs_external      init -inf */
  real *rbm;    /* sums_rightmost_basepairs_multibranch   init -inf */
  real *close;  /* sums_close       (sparse -> dense, -inf = absent) */
  real *acc;    /* sums_accessible  (sparse -> dense, -inf = absent) */
  real *mb;     /* sums_multibranch                       init -inf */
  real *m1;     /* sums_1ormore_basepairs                 init -inf */
  real *mbc;    /* fold_scores.multibranch_close_scores (memo, :333-335) */
  real *hps;    /* fold_scores.hairpin_scores     (-inf = key absent)   */
  real *accs;   /* fold_scores.accessible_scores  (-inf = key absent)   */
  /* Transposed mirrors (xT[j][i] == x[i][j], written together with x) of the matrices the recurrences walk DOWN A
   * COLUMN (`[k][j]`, k running): the same values read from contiguous memory.  Layout only — no arithmetic differs. */
  real *rbeT, *rbmT, *m1T;
} FoldSums;

static real *alloc_mat(int L, real v) {
  size_t n = (size_t)L * (size_t)L;
  real *m = (real *)malloc(n * sizeof(real));
  for (size_t x = 0; x < n; x++) m[x] = v;
  return m;
}

static void fold_sums_new(FoldSums *f, int L) {
  f->L = L;
  f->ext = alloc_mat(L, (real)0.);
  f->rbe = alloc_mat(L, NEG_INF);
  f->rbm = alloc_mat(L, NEG_INF);
  f->close = alloc_mat(L, NEG_INF);
  f->acc = alloc_mat(L, NEG_INF);
  f->mb = alloc_mat(L, NEG_INF);
  f->m1 = alloc_mat(L, NEG_INF);
  f->mbc = alloc_mat(L, (real)0.);
  f->hps = alloc_mat(L, NEG_INF);
  f->accs = alloc_mat(L, NEG_INF);
  f->rbeT = alloc_mat(L, NEG_INF);
  f->rbmT = alloc_mat(L, NEG_INF);
  f->m1T = alloc_mat(L, NEG_INF);
}

static void fold_sums_free(FoldSums *f) {
  free(f->ext); free(f->rbe); free(f->rbm); free(f->close); free(f->acc); free(f->mb); free(f->m1);
  free(f->mbc); free(f->rbeT); free(f->rbmT); free(f->m1T); free(f->hps); free(f->accs);
}

#define AT(m, i, j) (m)[(size_t)(i) * (size_t)L + (size_t)(j)]

/* get_fold_sums: src/mccaskill_algo.rs:282-378 (one cell of the span loop; the loops are in get_fold_sums below) */
static void fold_sums_cell(void *vc, int i) {
  const SpanCtx *c = (const SpanCtx *)vc;
  const uint8_t *s = c->s;
  const int L = c->L, span = c->span;
  TT *t = c->tt;
  FoldSums *f = c->f;
  const int MINSPAN = t->min_span_hairpin_close, MAX2 = t->max_2loop_len;
  const real CNB = (real)t->coeff_num_branches;
  {
    {
      int j = i + span - 1;
      real sum = NEG_INF;
      if (j - i + 1 >= MINSPAN && canonical(s[i], s[j])) {
        real hs = hairpin_score(s, i, j, t);
        AT(f->hps, i, j) = hs;              /* fold_scores.hairpin_scores.insert, :299-301 */
        lse(&sum, hs);
        for (int k = i + 1; k < j - 1; k++) {
          if (k - i - 1 > MAX2) break;
          for (int l = j - 1; l > k; l--) {
            if (j - l - 1 + k - i - 1 > MAX2) break;
            real x = AT(f->close, k, l);
            if (x > NEG_INF) {
              real y = twoloop_score(s, i, j, k, l, t);
              y = x + y;
              lse(&sum, y);
            }
          }
        }
        real mbc = multibranch_close_score(s, i, j, t);
        lse(&sum, AT(f->mb, i + 1, j - 1) + mbc);
        real as = accessible_score(s, L, i, j, t);
        if (sum > NEG_INF) {
          AT(f->mbc, i, j) = mbc;
          AT(f->accs, i, j) = as;
          AT(f->close, i, j) = sum;
          AT(f->acc, i, j) = sum + as;
        }
      }
      sum = NEG_INF;
      for (int k = i + 1; k <= j; k++) {
        real x = AT(f->acc, i, k);
        if (x > NEG_INF) lse(&sum, x);
      }
      AT(f->rbe, i, j) = sum;
      AT(f->rbeT, j, i) = sum;
      sum = (real)0.;
      for (int k = i; k < j; k++) {
        real x = AT(f->rbeT, j, k);
        real y = (i == 0 && k == 0) ? (real)0. : AT(f->ext, i, k - 1);
        y = x + y;
        lse(&sum, y);
      }
      AT(f->ext, i, j) = sum;
      sum = AT(f->rbe, i, j) + CNB;
      real sum2 = NEG_INF;
      for (int k = i + 1; k < j; k++) {
        real x = AT(f->rbeT, j, k) + CNB;
        lse(&sum, x);
        real y = AT(f->m1, i, k - 1) + x;
        lse(&sum2, y);
      }
      AT(f->mb, i, j) = sum2;
      lse(&sum, sum2);
      AT(f->m1, i, j) = sum;
      AT(f->m1T, j, i) = sum;
    }
  }
}
static void get_fold_sums(const uint8_t *s, int L, TT *t, FoldSums *f) {
  SpanCtx c;
  memset(&c, 0, sizeof c);
  c.s = s; c.L = L; c.tt = t; c.f = f;
  for (int span = t->min_span_hairpin_close; span <= L; span++) {   /* i ascending within a span, :283-284 */
    c.span = span;
    span_for(L - span + 1, fold_sums_cell, &c);
  }
}

/* get_fold_sums_contra: src/mccaskill_algo.rs:380-516 (one cell) */
static void fold_sums_contra_cell(void *vc, int i) {
  const SpanCtx *c = (const SpanCtx *)vc;
  const uint8_t *s = c->s;
  const int L = c->L, span = c->span, allows_short = c->allows_short;
  CT *t = c->ct;
  FoldSums *f = c->f;
  const int MINSPAN = t->min_span_hairpin_close, MAXL = t->max_loop_len;
  {
    {
      int j = i + span - 1;
      real sum = NEG_INF;
      if (canonical(s[i], s[j]) && (allows_short || j - i + 1 >= MINSPAN)) {
        if (j - i - 1 <= MAXL) {
          real hs = c_hairpin_score(s, i, j, t);
          AT(f->hps, i, j) = hs;            /* :408-410 */
          lse(&sum, hs);
        }
        for (int k = i + 1; k < j - 1; k++) {
          if (k - i - 1 > MAXL) break;
          for (int l = j - 1; l > k; l--) {
            if (j - l - 1 + k - i - 1 > MAXL) break;
            real x = AT(f->close, k, l);
            if (x > NEG_INF) {
              real y = c_twoloop_score(s, i, j, k, l, t);
              y = x + y;
              lse(&sum, y);
            }
          }
        }
        real mbc = (real)t->multibranch_score_base + (real)t->multibranch_score_basepair +
                   c_junction(s, L, i, j, t);
        lse(&sum, AT(f->mb, i + 1, j - 1) + mbc);
        real as = c_junction(s, L, j, i, t) + (real)t->basepair_scores[s[i]][s[j]];
        if (sum > NEG_INF) {
          AT(f->mbc, i, j) = mbc;
          AT(f->accs, i, j) = as;
          AT(f->close, i, j) = sum;
          AT(f->acc, i, j) = sum + as;
        }
      }
      sum = NEG_INF;
      real sum2 = sum;
      for (int k = i + 1; k <= j; k++) {
        real x = AT(f->acc, i, k);
        if (x > NEG_INF) {
          lse(&sum, x + (real)t->external_score_basepair + (real)t->external_score_unpair * (real)(j - k));
          lse(&sum2,
              x + (real)t->multibranch_score_basepair + (real)t->multibranch_score_unpair * (real)(j - k));
        }
      }
      AT(f->rbe, i, j) = sum;
      AT(f->rbeT, j, i) = sum;
      AT(f->rbm, i, j) = sum2;
      AT(f->rbmT, j, i) = sum2;
      sum = (real)t->external_score_unpair * (real)span;
      for (int k = i; k < j; k++) {
        real x = AT(f->rbeT, j, k);
        real y = (i == 0 && k == 0) ? (real)0. : AT(f->ext, i, k - 1);
        y = x + y;
        lse(&sum, y);
      }
      AT(f->ext, i, j) = sum;
      sum = AT(f->rbm, i, j);
      sum2 = NEG_INF;
      for (int k = i + 1; k < j; k++) {
        real x = AT(f->rbmT, j, k);
        lse(&sum, x + (real)t->multibranch_score_unpair * (real)(k - i));
        real y = AT(f->m1, i, k - 1) + x;
        lse(&sum2, y);
      }
      AT(f->mb, i, j) = sum2;
      lse(&sum, sum2);
      AT(f->m1, i, j) = sum;
      AT(f->m1T, j, i) = sum;
    }
  }
}
static void get_fold_sums_contra(const uint8_t *s, int L, int allows_short, CT *t, FoldSums *f) {
  SpanCtx c;
  memset(&c, 0, sizeof c);
  c.s = s; c.L = L; c.ct = t; c.f = f; c.allows_short = allows_short;
  for (int span = 1; span <= L; span++) {
    c.span = span;
    span_for(L - span + 1, fold_sums_contra_cell, &c);
  }
}

/* get_basepair_probs: src/mccaskill_algo.rs:518-610 (one cell).  Log-probs go to `bp` (dense, -inf = absent). */
static void basepair_probs_cell(void *vc, int i) {
  const SpanCtx *c = (const SpanCtx *)vc;
  const uint8_t *s = c->s;
  const int L = c->L, span = c->span;
  TT *t = c->tt;
  const FoldSums *f = c->f;
  real *bp = (real *)c->bp, *pm = (real *)c->pm, *pm2 = (real *)c->pm2;
  const int MAX2 = t->max_2loop_len;
  const real CNB = (real)t->coeff_num_branches;
  const real global_sum = (real)c->global_sum;
  {
    {
      int j = i + span - 1;
      real sum = NEG_INF, sum2 = NEG_INF;
      for (int k = j + 1; k < L; k++) {
        real x = AT(f->close, i, k);
        if (x > NEG_INF) {
          real p = AT(bp, i, k);
          real mbc = AT(f->mbc, i, k);
          x = p + mbc - x;
          lse(&sum, x + AT(f->m1, j + 1, k - 1));
          lse(&sum2, x);
        }
      }
      AT(pm, j, i) = sum;     /* probs_multibranch[i][j], stored transposed: read as [k][j], k running */
      AT(pm2, j, i) = sum2;
      real sum_close = AT(f->close, i, j);
      if (sum_close > NEG_INF) {
        real sum_acc = AT(f->acc, i, j);
        real sp0 = i < 1 ? (real)0. : AT(f->ext, 0, i - 1);
        real sp1 = j > L - 2 ? (real)0. : AT(f->ext, j + 1, L - 1);
        real sm = sp0 + sum_acc + sp1 - global_sum;
        for (int k = i - 1; k >= 0; k--) {
          if (i - k - 1 > MAX2) break;
          for (int l = j + 1; l < L; l++) {
            if (l - j - 1 + i - k - 1 > MAX2) break;
            real x = AT(f->close, k, l);
            if (x > NEG_INF) {
              lse(&sm, AT(bp, k, l) + sum_close - x + twoloop_score(s, k, l, i, j, t));
            }
          }
        }
        sum_acc = sum_acc + CNB;
        for (int k = 0; k < i; k++) {
          real x = AT(f->m1T, i - 1, k + 1);
          lse(&sm, sum_acc + AT(pm2, j, k) + x);
          real y = AT(pm, j, k);
          lse(&sm, sum_acc + y);
          lse(&sm, sum_acc + x + y);
        }
        if (sm > NEG_INF) AT(bp, i, j) = sm;
      }
    }
  }
}
static void get_basepair_probs(const uint8_t *s, int L, TT *t, const FoldSums *f, real *bp) {
  const int MINSPAN = t->min_span_hairpin_close;
  real *pm = alloc_mat(L, NEG_INF), *pm2 = alloc_mat(L, NEG_INF);
  SpanCtx c;
  memset(&c, 0, sizeof c);
  c.s = s; c.L = L; c.tt = t; c.f = (FoldSums *)f; c.bp = bp; c.pm = pm; c.pm2 = pm2;
  c.global_sum = (double)AT(f->ext, 0, L - 1);
  for (int span = L; span >= MINSPAN; span--) {
    c.span = span;
    span_for(L - span + 1, basepair_probs_cell, &c);
  }
  free(pm);
  free(pm2);
}

/* get_basepair_probs_contra: src/mccaskill_algo.rs:612-723 (one cell) */
static void basepair_probs_contra_cell(void *vc, int i) {
  const SpanCtx *c = (const SpanCtx *)vc;
  const uint8_t *s = c->s;
  const int L = c->L, span = c->span;
  CT *t = c->ct;
  const FoldSums *f = c->f;
  real *bp = (real *)c->bp, *pm = (real *)c->pm, *pm2 = (real *)c->pm2;
  const int MAXL = t->max_loop_len;
  const real global_sum = (real)c->global_sum;
  {
    {
      int j = i + span - 1;
      real sum = NEG_INF, sum2 = NEG_INF;
      for (int k = j + 1; k < L; k++) {
        real x = AT(f->close, i, k);
        if (x > NEG_INF) {
          real p = AT(bp, i, k);
          real mbc = AT(f->mbc, i, k);
          x = p + mbc - x;
          lse(&sum, x + AT(f->m1, j + 1, k - 1));
          lse(&sum2, x + (real)t->multibranch_score_unpair * (real)(k - j - 1));
        }
      }
      AT(pm, j, i) = sum;     /* probs_multibranch[i][j], stored transposed: read as [k][j], k running */
      AT(pm2, j, i) = sum2;
      real sum_close = AT(f->close, i, j);
      if (sum_close > NEG_INF) {
        real sp0 = i < 1 ? (real)0. : AT(f->ext, 0, i - 1);
        real sp1 = j > L - 2 ? (real)0. : AT(f->ext, j + 1, L - 1);
        real sm = sp0 + sp1 + AT(f->acc, i, j) + (real)t->external_score_basepair - global_sum;
        for (int k = i - 1; k >= 0; k--) {
          if (i - k - 1 > MAXL) break;
          for (int l = j + 1; l < L; l++) {
            if (l - j - 1 + i - k - 1 > MAXL) break;
            real x = AT(f->close, k, l);
            if (x > NEG_INF) {
              lse(&sm, AT(bp, k, l) + sum_close - x + c_twoloop_score(s, k, l, i, j, t));
            }
          }
        }
        real sum_acc = AT(f->acc, i, j) + (real)t->multibranch_score_basepair;
        for (int k = 0; k < i; k++) {
          real x = AT(f->m1T, i - 1, k + 1);
          lse(&sm, sum_acc + AT(pm2, j, k) + x);
          real y = AT(pm, j, k);
          lse(&sm, sum_acc + y + (real)t->multibranch_score_unpair * (real)(i - k - 1));
          lse(&sm, sum_acc + x + y);
        }
        if (sm > NEG_INF) AT(bp, i, j) = sm;
      }
    }
  }
}
static void get_basepair_probs_contra(const uint8_t *s, int L, int allows_short, CT *t,
                                      const FoldSums *f, real *bp) {
  const int MINSPAN = allows_short ? 2 : t->min_span_hairpin_close;
  real *pm = alloc_mat(L, NEG_INF), *pm2 = alloc_mat(L, NEG_INF);
  SpanCtx c;
  memset(&c, 0, sizeof c);
  c.s = s; c.L = L; c.ct = t; c.f = (FoldSums *)f; c.bp = bp; c.pm = pm; c.pm2 = pm2; c.allows_short = allows_short;
  c.global_sum = (double)AT(f->ext, 0, L - 1);
  for (int span = L; span >= MINSPAN; span--) {
    c.span = span;
    span_for(L - span + 1, basepair_probs_contra_cell, &c);
  }
  free(pm);
  free(pm2);
}

/* ---------------------------------------------------------------------------------------------
 * mccaskill_algo: src/mccaskill_algo.rs:247-280
 *   out_bpp: packed rna_bpp_len(L) floats (RNA_BPP_ABSENT for missing keys), may be NULL
 *   dbg_*  : optional dense L x L dumps (float), -inf = absent / never written
 * ------------------------------------------------------------------------------------------- */
int orc_mccaskill_algo(const uint8_t *seq, int L, int uses_contra_model, int allows_short_hairpins,
                       const RnaTurnerTables *tt, const RnaContraTables *ct, float *out_bpp,
                       float *out_logz, float *dbg_close, float *dbg_external, float *dbg_logprob,
                       float *dbg_m1) {
  if (L < 1) return RNA_ERR_EMPTY_SEQ;
  FoldSums f;
  fold_sums_new(&f, L);
  real *bp = alloc_mat(L, NEG_INF);
  if (uses_contra_model) {
    get_fold_sums_contra(seq, L, allows_short_hairpins, ct, &f);
    get_basepair_probs_contra(seq, L, allows_short_hairpins, ct, &f, bp);
  } else {
    get_fold_sums(seq, L, tt, &f);
    get_basepair_probs(seq, L, tt, &f, bp);
  }
  if (out_logz) *out_logz = (float)AT(f.ext, 0, L - 1);
  if (out_bpp) {
    for (int i = 0; i < L; i++)
      for (int j = i + 1; j < L; j++) {
        real v = AT(bp, i, j);
        /* basepair_probs.iter().map(|(x, &y)| (*x, expf(y))): src/mccaskill_algo.rs:608,721 */
        out_bpp[rna_bpp_index((uint64_t)L, (uint64_t)i, (uint64_t)j)] =
            v > NEG_INF ? (float)approx_expf(v) : RNA_BPP_ABSENT;
      }
  }
  size_t n = (size_t)L * (size_t)L;
  if (dbg_close) for (size_t x = 0; x < n; x++) dbg_close[x] = (float)f.close[x];
  if (dbg_external) for (size_t x = 0; x < n; x++) dbg_external[x] = (float)f.ext[x];
  if (dbg_logprob) for (size_t x = 0; x < n; x++) dbg_logprob[x] = (float)bp[x];
  if (dbg_m1) for (size_t x = 0; x < n; x++) dbg_m1[x] = (float)f.m1[x];
  free(bp);
  fold_sums_free(&f);
  return RNA_OK;
}

/* the max-plus fill of centroid_fold, src/centroid_fold.rs:33-64 (one cell) */
typedef struct { const float *bpp; int L, span; real g; real *W, *WT; } CentroidCtx;
#define HAS(i, j) (bpp[rna_bpp_index((uint64_t)L, (uint64_t)(i), (uint64_t)(j))] != RNA_BPP_ABSENT)
#define PR(i, j) ((real)bpp[rna_bpp_index((uint64_t)L, (uint64_t)(i), (uint64_t)(j))])
static void centroid_cell(void *vc, int i) {
  const CentroidCtx *c = (const CentroidCtx *)vc;
  const float *bpp = c->bpp;
  const int L = c->L, span = c->span;
  const real g = c->g;
  real *W = c->W, *WT = c->WT;   /* WT[j][i] == W[i][j]: the split loop reads W[k+1][j] down a column */
  {
    {
      int j = i + span - 1;
      if (i == j) return;
      real w = AT(W, i + 1, j);
      real e = AT(W, i, j - 1);
      if (e > w) w = e;
      if (HAS(i, j)) {
        e = AT(W, i + 1, j - 1) + g * PR(i, j) - (real)1.;
        if (e > w) w = e;
      }
      for (int k = i + 1; k < j; k++) {
        e = AT(W, i, k) + AT(WT, j, k + 1);
        if (e > w) w = e;
      }
      AT(W, i, j) = w;
      AT(WT, j, i) = w;
    }
  }
}
#undef HAS
#undef PR

/* get_fold_sums / get_fold_sums_contra stand-alone + the FoldScores memo (src/mccaskill_algo.rs:3-22, 282-516), as dense
 * L x L planes in the order of include/rna_algos_b200.h RNA_SUMS_*: -inf = key absent (hash-map members). */
int orc_fold_sums(const uint8_t *seq, int L, int uses_contra_model, int allows_short_hairpins,
                  const RnaTurnerTables *tt, const RnaContraTables *ct, float *planes) {
  if (L < 1) return RNA_ERR_EMPTY_SEQ;
  FoldSums f;
  fold_sums_new(&f, L);
  if (uses_contra_model) get_fold_sums_contra(seq, L, allows_short_hairpins, ct, &f);
  else get_fold_sums(seq, L, tt, &f);
  const size_t n = (size_t)L * (size_t)L;
  const real *src[RNA_SUMS_PLANES] = {f.close, f.acc, f.ext, f.rbe, f.rbm, f.mb, f.m1, f.hps, f.mbc, f.accs};
  for (int p = 0; p < RNA_SUMS_PLANES; p++)
    for (size_t x = 0; x < n; x++) {
      float v = (float)src[p][x];
      if (p == RNA_SCORES_MB_CLOSE && !(f.close[x] > NEG_INF)) v = -INFINITY;   /* inserted with sums_close only */
      planes[(size_t)p * n + x] = v;
    }
  fold_sums_free(&f);
  return RNA_OK;
}

/* ---------------------------------------------------------------------------------------------
 * centroid_fold: src/centroid_fold.rs:25-105.  `bpp` is the packed matrix (absent = RNA_BPP_ABSENT).
 * ------------------------------------------------------------------------------------------- */
int orc_centroid_fold(const float *bpp, int L, float centroid_threshold, uint8_t *out_fold_str,
                      uint16_t *out_pairs, uint32_t *out_num_pairs, float *out_expect_accuracy) {
  if (L < 1) return RNA_ERR_EMPTY_SEQ;
  real g = (real)centroid_threshold;
  real *W = alloc_mat(L, (real)0.), *WT = alloc_mat(L, (real)0.);
#define HAS(i, j) (bpp[rna_bpp_index((uint64_t)L, (uint64_t)(i), (uint64_t)(j))] != RNA_BPP_ABSENT)
#define PR(i, j) ((real)bpp[rna_bpp_index((uint64_t)L, (uint64_t)(i), (uint64_t)(j))])
  {
    CentroidCtx cc = {bpp, L, 0, g, W, WT};
    for (int span = 1; span <= L; span++) {
      cc.span = span;
      span_for(L - span + 1, centroid_cell, &cc);
    }
  }
  uint32_t np = 0;
  if (out_fold_str) memset(out_fold_str, '.', (size_t)L);
  int *stack = (int *)malloc(sizeof(int) * 2 * (size_t)(L + 2));
  int sp = 0;
  stack[0] = 0; stack[1] = L - 1; sp = 1;
  while (sp > 0) {
    sp--;
    int i = stack[2 * sp], j = stack[2 * sp + 1];
    if (j <= i) continue;
    real w = AT(W, i, j);
    if (w == (real)0.) continue;
    if (w == AT(W, i + 1, j)) {
      stack[2 * sp] = i + 1; stack[2 * sp + 1] = j; sp++;
    } else if (w == AT(W, i, j - 1)) {
      stack[2 * sp] = i; stack[2 * sp + 1] = j - 1; sp++;
    } else if (HAS(i, j) && w == AT(W, i + 1, j - 1) + g * PR(i, j) - (real)1.) {
      stack[2 * sp] = i + 1; stack[2 * sp + 1] = j - 1; sp++;
      if (out_pairs) { out_pairs[2 * np] = (uint16_t)i; out_pairs[2 * np + 1] = (uint16_t)j; }
      if (out_fold_str) { out_fold_str[i] = '('; out_fold_str[j] = ')'; }   /* get_fold_str, bin:197-207 */
      np++;
    } else {
      for (int k = i + 1; k < j; k++) {
        if (w == AT(W, i, k) + AT(W, k + 1, j)) {
          stack[2 * sp] = i; stack[2 * sp + 1] = k; sp++;
          stack[2 * sp] = k + 1; stack[2 * sp + 1] = j; sp++;
          break;
        }
      }
    }
  }
  if (out_num_pairs) *out_num_pairs = np;
  if (out_expect_accuracy) *out_expect_accuracy = (float)AT(W, 0, L - 1);
  free(stack);
  free(W);
  free(WT);
#undef HAS
#undef PR
  return RNA_OK;
}

/* ---------------------------------------------------------------------------------------------
 * Durbin pair-HMM: src/durbin_algo.rs:73-242.  s0/s1 are SENTINEL-PADDED (n = len+2).
 * ------------------------------------------------------------------------------------------- */
static void durbin_padded(const uint8_t *s0, int n, const uint8_t *s1, int m, const RnaAlignTables *a,
                          float *out) {
  size_t sz = (size_t)n * (size_t)m;
  real *fm = (real *)malloc(sz * sizeof(real)), *fi = (real *)malloc(sz * sizeof(real)),
       *fd = (real *)malloc(sz * sizeof(real)), *bm = (real *)malloc(sz * sizeof(real)),
       *bi = (real *)malloc(sz * sizeof(real)), *bd = (real *)malloc(sz * sizeof(real));
  for (size_t x = 0; x < sz; x++) fm[x] = fi[x] = fd[x] = bm[x] = bi[x] = bd[x] = NEG_INF;
#define D(mt, i, j) (mt)[(size_t)(i) * (size_t)m + (size_t)(j)]
  const real m2m = (real)a->match2match_score, m2i = (real)a->match2insert_score,
             iex = (real)a->insert_extend_score, inm = (real)a->init_match_score,
             ini = (real)a->init_insert_score;
  /* forward: src/durbin_algo.rs:82-139 */
  for (int i = 0; i < n - 1; i++) {
    for (int j = 0; j < m - 1; j++) {
      if (i == 0 && j == 0) { D(fm, i, j) = (real)0.; continue; }
      if (i > 0 && j > 0) {
        real sum = NEG_INF;
        real ms = (real)a->match_scores[s0[i]][s1[j]];
        int begins = (i - 1 == 0 && j - 1 == 0);
        lse(&sum, D(fm, i - 1, j - 1) + (begins ? inm : m2m));
        lse(&sum, D(fi, i - 1, j - 1) + m2i);
        lse(&sum, D(fd, i - 1, j - 1) + m2i);
        D(fm, i, j) = sum + ms;
      }
      if (i > 0) {
        real is = (real)a->insert_scores[s0[i]];
        int begins = (i - 1 == 0 && j == 0);
        real sum = NEG_INF;
        lse(&sum, D(fm, i - 1, j) + (begins ? ini : m2i));
        lse(&sum, D(fi, i - 1, j) + iex);
        D(fi, i, j) = sum + is;
      }
      if (j > 0) {
        real is = (real)a->insert_scores[s1[j]];
        int begins = (i == 0 && j - 1 == 0);
        real sum = NEG_INF;
        lse(&sum, D(fm, i, j - 1) + (begins ? ini : m2i));
        lse(&sum, D(fd, i, j - 1) + iex);
        D(fd, i, j) = sum + is;
      }
    }
  }
  /* backward: src/durbin_algo.rs:140-197 */
  for (int i = n - 1; i >= 1; i--) {
    for (int j = m - 1; j >= 1; j--) {
      if (i == n - 1 && j == m - 1) { D(bm, i, j) = (real)0.; continue; }
      if (i < n - 1 && j < m - 1) {
        real sum = NEG_INF;
        real ms = (real)a->match_scores[s0[i]][s1[j]];
        int ends = (i + 1 == n - 1 && j + 1 == m - 1);
        lse(&sum, D(bm, i + 1, j + 1) + (ends ? (real)0. : m2m));
        lse(&sum, D(bi, i + 1, j + 1) + m2i);
        lse(&sum, D(bd, i + 1, j + 1) + m2i);
        D(bm, i, j) = sum + ms;
      }
      if (i < n - 1) {
        real is = (real)a->insert_scores[s0[i]];
        int ends = (i + 1 == n - 1 && j == m - 1);
        real sum = NEG_INF;
        lse(&sum, D(bm, i + 1, j) + (ends ? (real)0. : m2i));
        lse(&sum, D(bi, i + 1, j) + iex);
        D(bi, i, j) = sum + is;
      }
      if (j < m - 1) {
        real is = (real)a->insert_scores[s1[j]];
        int ends = (i == n - 1 && j + 1 == m - 1);
        real sum = NEG_INF;
        lse(&sum, D(bm, i, j + 1) + (ends ? (real)0. : m2i));
        lse(&sum, D(bd, i, j + 1) + iex);
        D(bd, i, j) = sum + is;
      }
    }
  }
  /* get_match_probs: src/durbin_algo.rs:201-242 */
  for (size_t x = 0; x < sz; x++) out[x] = 0.f;
  real global_sum = D(fm, n - 2, m - 2);
  lse(&global_sum, D(fi, n - 2, m - 2));
  lse(&global_sum, D(fd, n - 2, m - 2));
  for (int i = 1; i < n - 1; i++) {
    for (int j = 1; j < m - 1; j++) {
      real sum = NEG_INF;
      real fwd = D(fm, i, j);
      int ends = (i + 1 == n - 1 && j + 1 == m - 1);
      lse(&sum, (ends ? (real)0. : m2m) + D(bm, i + 1, j + 1));
      lse(&sum, m2i + D(bi, i + 1, j + 1));
      lse(&sum, m2i + D(bd, i + 1, j + 1));
      out[(size_t)i * (size_t)m + (size_t)j] = (float)approx_expf(fwd + sum - global_sum);
    }
  }
#undef D
  free(fm); free(fi); free(fd); free(bm); free(bi); free(bd);
}

/* durbin_algo on sentinel-FREE inputs; adds PSEUDO_BASE at both ends like src/bin/durbin_algo.rs:48-50.
 * out: (la+2) x (lb+2) dense row-major. */
int orc_durbin_algo(const uint8_t *sa, int la, const uint8_t *sb, int lb, const RnaAlignTables *a,
                    float *out) {
  if (la < 1 || lb < 1) return RNA_ERR_EMPTY_SEQ;
  int n = la + 2, m = lb + 2;
  uint8_t *p0 = (uint8_t *)malloc((size_t)n), *p1 = (uint8_t *)malloc((size_t)m);
  p0[0] = p0[n - 1] = RNA_PSEUDO_BASE;
  p1[0] = p1[m - 1] = RNA_PSEUDO_BASE;
  memcpy(p0 + 1, sa, (size_t)la);
  memcpy(p1 + 1, sb, (size_t)lb);
  durbin_padded(p0, n, p1, m, a, out);
  free(p0);
  free(p1);
  return RNA_OK;
}

/* FoldScoreSets::accumulate restated independently of the product: src/mccaskill_algo.rs:60-86 */
void orc_contra_accumulate(RnaContraTables *t) {
  float sum = 0.f;
  for (int i = 0; i < RNA_CONTRA_MAX_LOOP_LEN + 1; i++) { sum += t->hairpin_scores_len[i]; t->hairpin_scores_len_cumulative[i] = sum; }
  sum = 0.f;
  for (int i = 0; i < RNA_CONTRA_MAX_LOOP_LEN; i++) { sum += t->bulge_scores_len[i]; t->bulge_scores_len_cumulative[i] = sum; }
  sum = 0.f;
  for (int i = 0; i < RNA_CONTRA_MAX_LOOP_LEN - 1; i++) { sum += t->interior_scores_len[i]; t->interior_scores_len_cumulative[i] = sum; }
  sum = 0.f;
  for (int i = 0; i < RNA_CONTRA_MAX_INTERIOR_SYMMETRIC; i++) { sum += t->interior_scores_symmetric[i]; t->interior_scores_symmetric_cumulative[i] = sum; }
  sum = 0.f;
  for (int i = 0; i < RNA_CONTRA_MAX_INTERIOR_ASYMMETRIC; i++) { sum += t->interior_scores_asymmetric[i]; t->interior_scores_asymmetric_cumulative[i] = sum; }
}

/* Scorer probes for the brute-force enumerator in tests/ (model-energy of one loop). */
double orc_score_hairpin(const uint8_t *s, int L, int i, int j, int contra, const RnaTurnerTables *tt, const RnaContraTables *ct) {
  (void)L;
  return contra ? (double)c_hairpin_score(s, i, j, ct) : (double)hairpin_score(s, i, j, tt);
}
double orc_score_twoloop(const uint8_t *s, int L, int i, int j, int k, int l, int contra, const RnaTurnerTables *tt, const RnaContraTables *ct) {
  (void)L;
  return contra ? (double)c_twoloop_score(s, i, j, k, l, ct) : (double)twoloop_score(s, i, j, k, l, tt);
}
double orc_score_multibranch_close(const uint8_t *s, int L, int i, int j, int contra, const RnaTurnerTables *tt, const RnaContraTables *ct) {
  if (contra) return (double)((real)ct->multibranch_score_base + (real)ct->multibranch_score_basepair + c_junction(s, L, i, j, ct));
  return (double)multibranch_close_score(s, i, j, tt);
}
double orc_score_accessible(const uint8_t *s, int L, int i, int j, int contra, const RnaTurnerTables *tt, const RnaContraTables *ct) {
  if (contra) return (double)(c_junction(s, L, j, i, ct) + (real)ct->basepair_scores[s[i]][s[j]]);
  return (double)accessible_score(s, L, i, j, tt);
}

/* ---------------------------------------------------------------------------------------------
 * Batch drivers: one unit (sequence / pair) per task on a pool of host threads, mirroring
 * benches/benches.rs:24-41 and src/bin/centroid_fold.rs:119-161.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  const uint8_t *bases; const uint32_t *offsets; uint32_t n_seqs;
  int model, allows_short;
  const RnaTurnerTables *tt; const RnaContraTables *ct;
  const float *gammas; uint32_t n_gammas;
  float *out_logz, *out_bpp; const uint64_t *bpp_offsets;
  uint8_t *out_structs; float *out_ea;
  uint32_t total_len;
  volatile uint32_t next;
  uint64_t lse_terms;
  pthread_mutex_t mu;
} FoldJob;

static void *fold_worker(void *arg) {
  FoldJob *jb = (FoldJob *)arg;
  g_lse_terms = 0;
  for (;;) {
    uint32_t sidx = __atomic_fetch_add(&jb->next, 1u, __ATOMIC_RELAXED);
    if (sidx >= jb->n_seqs) break;
    const uint8_t *seq = jb->bases + jb->offsets[sidx];
    int L = (int)(jb->offsets[sidx + 1] - jb->offsets[sidx]);
    float *bpp = NULL, *tmp = NULL;
    if (jb->out_bpp) bpp = jb->out_bpp + jb->bpp_offsets[sidx];
    else if (jb->n_gammas) bpp = tmp = (float *)malloc(sizeof(float) * (size_t)rna_bpp_len((uint64_t)L) + 4);
    orc_mccaskill_algo(seq, L, jb->model == RNA_MODEL_CONTRA, jb->allows_short, jb->tt, jb->ct, bpp,
                       jb->out_logz ? jb->out_logz + sidx : NULL, NULL, NULL, NULL, NULL);
    for (uint32_t g = 0; g < jb->n_gammas; g++) {
      orc_centroid_fold(bpp, L, jb->gammas[g],
                        jb->out_structs ? jb->out_structs + (size_t)g * jb->total_len + jb->offsets[sidx] : NULL,
                        NULL, NULL, jb->out_ea ? jb->out_ea + (size_t)g * jb->n_seqs + sidx : NULL);
    }
    free(tmp);
  }
  pthread_mutex_lock(&jb->mu);
  jb->lse_terms += g_lse_terms;
  pthread_mutex_unlock(&jb->mu);
  return NULL;
}

/* Returns the number of LSE-terms executed (algorithmic work counter) through *out_lse_terms. */
int orc_mccaskill_centroid_batch(const uint8_t *bases, const uint32_t *offsets, uint32_t n_seqs, int model,
                                 int allows_short, const RnaTurnerTables *tt, const RnaContraTables *ct,
                                 const float *gammas, uint32_t n_gammas, float *out_logz, float *out_bpp,
                                 const uint64_t *bpp_offsets, uint8_t *out_structs, float *out_ea,
                                 int n_threads, uint64_t *out_lse_terms) {
  FoldJob jb;
  memset(&jb, 0, sizeof jb);
  jb.bases = bases; jb.offsets = offsets; jb.n_seqs = n_seqs; jb.model = model; jb.allows_short = allows_short;
  jb.tt = tt; jb.ct = ct; jb.gammas = gammas; jb.n_gammas = n_gammas; jb.out_logz = out_logz;
  jb.out_bpp = out_bpp; jb.bpp_offsets = bpp_offsets; jb.out_structs = out_structs; jb.out_ea = out_ea;
  jb.total_len = offsets[n_seqs];
  pthread_mutex_init(&jb.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  if (g_inner_threads > 1) n_threads = 1;   /* the span team is one per process: sequences one after the other */
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (int x = 0; x < n_threads; x++) pthread_create(&th[x], NULL, fold_worker, &jb);
  for (int x = 0; x < n_threads; x++) pthread_join(th[x], NULL);
  free(th);
  if (out_lse_terms) *out_lse_terms = jb.lse_terms;
  return RNA_OK;
}

typedef struct {
  const uint8_t *bases; const uint32_t *offsets; const uint32_t *pairs; uint32_t n_pairs;
  const RnaAlignTables *a; float *out; const uint64_t *prob_offsets;
  volatile uint32_t next; uint64_t lse_terms; pthread_mutex_t mu;
} DurbinJob;

static void *durbin_worker(void *arg) {
  DurbinJob *jb = (DurbinJob *)arg;
  g_lse_terms = 0;
  for (;;) {
    uint32_t p = __atomic_fetch_add(&jb->next, 1u, __ATOMIC_RELAXED);
    if (p >= jb->n_pairs) break;
    uint32_t a = jb->pairs[2 * p], b = jb->pairs[2 * p + 1];
    orc_durbin_algo(jb->bases + jb->offsets[a], (int)(jb->offsets[a + 1] - jb->offsets[a]),
                    jb->bases + jb->offsets[b], (int)(jb->offsets[b + 1] - jb->offsets[b]), jb->a,
                    jb->out + jb->prob_offsets[p]);
  }
  pthread_mutex_lock(&jb->mu);
  jb->lse_terms += g_lse_terms;
  pthread_mutex_unlock(&jb->mu);
  return NULL;
}

int orc_durbin_batch(const uint8_t *bases, const uint32_t *offsets, const uint32_t *pairs, uint32_t n_pairs,
                     const RnaAlignTables *a, float *out, const uint64_t *prob_offsets, int n_threads,
                     uint64_t *out_lse_terms) {
  DurbinJob jb;
  memset(&jb, 0, sizeof jb);
  jb.bases = bases; jb.offsets = offsets; jb.pairs = pairs; jb.n_pairs = n_pairs; jb.a = a; jb.out = out;
  jb.prob_offsets = prob_offsets;
  pthread_mutex_init(&jb.mu, NULL);
  if (n_threads < 1) n_threads = 1;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
  for (int x = 0; x < n_threads; x++) pthread_create(&th[x], NULL, durbin_worker, &jb);
  for (int x = 0; x < n_threads; x++) pthread_join(th[x], NULL);
  free(th);
  if (out_lse_terms) *out_lse_terms = jb.lse_terms;
  return RNA_OK;
}

int orc_is_exact_flavour(void) {
#ifdef ORC_EXACT
  return 1;
#else
  return 0;
#endif
}
