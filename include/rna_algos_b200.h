/*
 * rna_algos_b200.h — C ABI of the B200-native McCaskill / centroid / Durbin hot path.
 *
 * This header is the drop-in boundary: every entry point below replaces one public
 * function of the reference crate `rna_algos` (heartsh/rna-algos 0.1.37).  Citations are
 * `path:line` relative to the reference tree.
 *
 *   reference (Rust)                                             this ABI
 *   ------------------------------------------------------------ -----------------------------------
 *   mccaskill_algo<T>()          src/mccaskill_algo.rs:247-280   rna_mccaskill_algo / rna_mccaskill_batch
 *   centroid_fold<T>()           src/centroid_fold.rs:25-105     rna_centroid_fold  / rna_centroid_batch
 *   bin centroid_fold (BPP+MEA)  src/bin/centroid_fold.rs:104-161 rna_mccaskill_centroid_batch (fused)
 *   durbin_algo()                src/durbin_algo.rs:73-77        rna_durbin_algo    / rna_durbin_batch
 *   FoldScoreSets (+transfer)    src/utils.rs:91-119,
 *                                src/mccaskill_algo.rs:24-211    RnaContraTables, rna_contra_tables_accumulate
 *   rna-ss-params Turner consts  src/utils.rs:8-10 (glob import) RnaTurnerTables
 *   AlignScores (+transfer)      src/durbin_algo.rs:4-57,
 *                                src/compiled_align_scores.rs    RnaAlignTables
 *
 * Plain C: pointers and sizes only, no C++/torch types.  All functions return an RNA_* status code
 * and never abort the process (the reference panics instead: src/utils.rs:570-572).
 *
 * Numerics: all DP values are f32, evaluated in the reference's order with the reference's
 * piecewise-cubic logsumexp / expf (src/utils.rs:579-655) and without FMA contraction, so results are
 * bit-identical to the reference algorithm given identical tables.
 */
#ifndef RNA_ALGOS_B200_H
#define RNA_ALGOS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------
 * Status codes
 * ---------------------------------------------------------------------------------------------- */
enum {
  RNA_OK = 0,
  RNA_ERR_BAD_ARG = 1,       /* null pointer, inconsistent offsets, bad enum                       */
  RNA_ERR_INVALID_BASE = 2,  /* a base code outside 0..3 (reference: bytes2seq panic)              */
  RNA_ERR_EMPTY_SEQ = 3,     /* L == 0 (reference underflows seq_len - 1, src/mccaskill_algo.rs:526) */
  RNA_ERR_TOO_LONG = 4,      /* L > RNA_MAX_SEQ_LEN (u16 index domain of the reference), or > RNA_MAX_FOLD_LEN for folding */
  RNA_ERR_NO_TABLES = 5,     /* the table blob for the requested model was never set               */
  RNA_ERR_BAD_TABLES = 6,    /* caps in the blob exceed the fixed array sizes below                */
  RNA_ERR_CUDA = 7,          /* a CUDA runtime call failed; see rna_last_error()                   */
  RNA_ERR_NO_DEVICE = 8,     /* no usable CUDA device: there is NO CPU fallback                    */
  RNA_ERR_NOMEM = 9
};

enum { RNA_MODEL_TURNER = 0, RNA_MODEL_CONTRA = 1 };

/* Numeric modes of the McCaskill passes (rna_set_numeric_mode).
 *   RNA_NUMERIC_REF_EXACT  (default) every fold in the reference's order with its piecewise-cubic logsumexp / expf
 *                          (src/utils.rs:579-655): results bit-identical to the reference algorithm.
 *   RNA_NUMERIC_FAST_F32   the speed mode: the same recurrences in exact log-space arithmetic (ex2.approx /
 *                          lg2.approx, f32 state).  The few long sequences of a call run one at a time on a cooperative
 *                          grid with warp-shuffle max / sum-of-exp reductions (3-8x faster than REF_EXACT at 1-4 k nt);
 *                          everything else runs on a second build of the batch kernel whose chains keep linear-space
 *                          sums (1.1x REF_EXACT on tRNA-length batches).
 *   RNA_NUMERIC_FAST_F64   the accuracy mode: f64 state and libdevice exp / log, warp-shuffle reductions throughout.
 * The FAST modes do NOT reproduce the reference's approximation error (<= 7.6e-6 per logsumexp): they agree with an
 * exact-math evaluation to the tolerances of DESIGN.md §2 and with the reference only to ~1e-3 in a probability. */
enum { RNA_NUMERIC_REF_EXACT = 0, RNA_NUMERIC_FAST_F32 = 1, RNA_NUMERIC_FAST_F64 = 2 };

#define RNA_NUM_BASES 4
#define RNA_BASE_A 0
#define RNA_BASE_C 1
#define RNA_BASE_G 2
#define RNA_BASE_U 3
#define RNA_PSEUDO_BASE 4        /* src/utils.rs:122 — only ever at Durbin sentinel positions       */
#define RNA_MAX_SEQ_LEN 65535u   /* u16 HashIndex domain, src/bin/centroid_fold.rs:85-101 (Durbin inputs) */
#define RNA_MAX_FOLD_LEN 46340   /* mccaskill / centroid inputs: triangular matrices are indexed with 32-bit
                                    offsets (46340^2 < 2^31); longer sequences get RNA_ERR_TOO_LONG from
                                    rna_validate_fold_lengths and from every fold entry point, before any copy */
#define RNA_LOOP_TABLE_LEN 31    /* lengths 0..30                                                   */
#define RNA_MAX_SPECIAL_HAIRPINS 128
#define RNA_MAX_SPECIAL_HAIRPIN_LEN 12
#define RNA_BPP_ABSENT (-1.0f)   /* "key not in SparseProbMat"; present-with-0.0 stays 0.0          */

/* ------------------------------------------------------------------------------------------------
 * Table blobs.  They are RUNTIME arguments (like &FoldScoreSets / &AlignScores in the reference), so
 * the genuine rna-ss-params values can be injected by a Rust shim without rebuilding the kernels.
 * Field names follow the symbols the reference consumes.  All scores are
 * dimensionless log-weights that are summed and fed to logsumexp directly.
 * ---------------------------------------------------------------------------------------------- */

/* One entry of HAIRPIN_SCORES_SPECIAL (src/utils.rs:198-205): the whole hairpin slice including the
 * closing pair, as base codes, and the score returned when the slice matches. */
typedef struct {
  uint8_t len;
  uint8_t seq[RNA_MAX_SPECIAL_HAIRPIN_LEN];
  uint8_t _pad[3];
  float score;
} RnaSpecialHairpin;

/* Turner 2004 model: the rna_ss_params::compiled_scores_turner consts + the model caps. */
typedef struct {
  int32_t max_2loop_len;                  /* MAX_2LOOP_LEN            (must be <= 30)              */
  int32_t min_span_hairpin_close;         /* MIN_SPAN_HAIRPIN_CLOSE                               */
  int32_t min_hairpin_len;                /* MIN_HAIRPIN_LEN                                      */
  int32_t max_hairpin_len_extrapolation;  /* MAX_HAIRPIN_LEN_EXTRAPOLATION (table used up to here) */
  int32_t min_hairpin_len_extrapolation;  /* MIN_HAIRPIN_LEN_EXTRAPOLATION                        */
  int32_t num_special_hairpins;
  float coeff_hairpin_len_extrapolation;  /* COEFF_HAIRPIN_LEN_EXTRAPOLATION                      */
  float helix_augu_end_penalty;           /* HELIX_AUGU_END_PENALTY                               */
  float ninio_coeff;                      /* NINIO_COEFF                                          */
  float ninio_max;                        /* NINIO_MAX                                            */
  float init_multibranch_base;            /* INIT_MULTIBRANCH_BASE                                */
  float coeff_num_branches;               /* COEFF_NUM_BRANCHES                                   */
  float hairpin_scores_init[RNA_LOOP_TABLE_LEN];   /* HAIRPIN_SCORES_INIT[len]                    */
  float bulge_scores_init[RNA_LOOP_TABLE_LEN];     /* BULGE_SCORES_INIT[len]                      */
  float interior_scores_init[RNA_LOOP_TABLE_LEN];  /* INTERIOR_SCORES_INIT[len]                   */
  float stack_scores[4][4][4][4];                        /* STACK_SCORES[i][j][k][l]              */
  float terminal_mismatch_scores_hairpin[4][4][4][4];    /* [i][j][i+1][j-1]                      */
  float terminal_mismatch_scores_1xmany[4][4][4][4];
  float terminal_mismatch_scores_2x3[4][4][4][4];
  float terminal_mismatch_scores_interior[4][4][4][4];
  float terminal_mismatch_scores_multibranch[4][4][4][4];
  float dangling_scores_5prime[4][4][4];                 /* [i][j][i-1]                           */
  float dangling_scores_3prime[4][4][4];                 /* [i][j][j+1]                           */
  float interior_scores_1x1[4][4][4][4][4][4];           /* src/utils.rs:275-276                  */
  float interior_scores_1x2[4][4][4][4][4][4][4];        /* src/utils.rs:283-284                  */
  float interior_scores_2x2[4][4][4][4][4][4][4][4];     /* src/utils.rs:302-303                  */
  RnaSpecialHairpin hairpin_scores_special[RNA_MAX_SPECIAL_HAIRPINS];
} RnaTurnerTables;

#define RNA_CONTRA_MAX_LOOP_LEN 30
#define RNA_CONTRA_MAX_INTERIOR_SYMMETRIC 15
#define RNA_CONTRA_MAX_INTERIOR_ASYMMETRIC 28
#define RNA_CONTRA_MAX_INTERIOR_EXPLICIT 4

/* CONTRAfold v2.02 model: a field-for-field image of FoldScoreSets (src/utils.rs:91-119) plus the
 * caps the recurrences read.  The *_cumulative arrays are what the scorers use; fill them with
 * rna_contra_tables_accumulate() (mirrors FoldScoreSets::accumulate, src/mccaskill_algo.rs:60-86). */
typedef struct {
  int32_t max_loop_len;             /* MAX_LOOP_LEN            (must be == 30 array bound)         */
  int32_t min_span_hairpin_close;   /* MIN_SPAN_HAIRPIN_CLOSE                                      */
  int32_t max_interior_explicit;    /* MAX_INTERIOR_EXPLICIT   (must be <= 4)                      */
  int32_t _pad;
  float hairpin_scores_len[RNA_CONTRA_MAX_LOOP_LEN + 1];
  float bulge_scores_len[RNA_CONTRA_MAX_LOOP_LEN];
  float interior_scores_len[RNA_CONTRA_MAX_LOOP_LEN - 1];
  float interior_scores_symmetric[RNA_CONTRA_MAX_INTERIOR_SYMMETRIC];
  float interior_scores_asymmetric[RNA_CONTRA_MAX_INTERIOR_ASYMMETRIC];
  float stack_scores[4][4][4][4];
  float terminal_mismatch_scores[4][4][4][4];
  float dangling_scores_left[4][4][4];
  float dangling_scores_right[4][4][4];
  float helix_close_scores[4][4];
  float basepair_scores[4][4];
  float interior_scores_explicit[RNA_CONTRA_MAX_INTERIOR_EXPLICIT][RNA_CONTRA_MAX_INTERIOR_EXPLICIT];
  float bulge_scores_0x1[4];
  float interior_scores_1x1[4][4];
  float multibranch_score_base;
  float multibranch_score_basepair;
  float multibranch_score_unpair;
  float external_score_basepair;
  float external_score_unpair;
  float hairpin_scores_len_cumulative[RNA_CONTRA_MAX_LOOP_LEN + 1];
  float bulge_scores_len_cumulative[RNA_CONTRA_MAX_LOOP_LEN];
  float interior_scores_len_cumulative[RNA_CONTRA_MAX_LOOP_LEN - 1];
  float interior_scores_symmetric_cumulative[RNA_CONTRA_MAX_INTERIOR_SYMMETRIC];
  float interior_scores_asymmetric_cumulative[RNA_CONTRA_MAX_INTERIOR_ASYMMETRIC];
} RnaContraTables;

/* CONTRAlign v2.01 pair-HMM scores: image of AlignScores (src/durbin_algo.rs:4-14). */
typedef struct {
  float match2match_score;
  float match2insert_score;
  float insert_extend_score;
  float insert_switch_score;   /* loaded but never read by the reference           */
  float init_match_score;
  float init_insert_score;
  float insert_scores[4];
  float match_scores[4][4];
} RnaAlignTables;

/* Sequential f32 prefix sums of the five *_len / symmetric / asymmetric arrays into the
 * *_cumulative arrays (src/mccaskill_algo.rs:60-86).  Pure host helper. */
void rna_contra_tables_accumulate(RnaContraTables *t);

/* The in-tree CONTRAlign v2.01 constants (src/compiled_align_scores.rs:2-19). */
void rna_align_tables_contralign_v201(RnaAlignTables *t);

/* ------------------------------------------------------------------------------------------------
 * Handle: one per GPU.  Thread-compatible (one in-flight call per handle; any number of handles).
 * ---------------------------------------------------------------------------------------------- */
typedef struct rna_handle rna_handle;

int rna_create(int device, rna_handle **out);       /* RNA_ERR_NO_DEVICE if CUDA is unusable       */
int rna_destroy(rna_handle *h);
const char *rna_last_error(const rna_handle *h);    /* human-readable detail of the last failure   */
int rna_device(const rna_handle *h);

int rna_set_numeric_mode(rna_handle *h, int mode);  /* RNA_NUMERIC_*; applies to the mccaskill entry points   */
int rna_get_numeric_mode(const rna_handle *h);

int rna_set_turner_tables(rna_handle *h, const RnaTurnerTables *t);
int rna_set_contra_tables(rna_handle *h, const RnaContraTables *t);
int rna_set_align_tables(rna_handle *h, const RnaAlignTables *t);

/* Number of f32 slots of one sequence's packed BPP matrix: entries (i,j), i<j, row-major:
 *   index(i,j) = i*(2L-i-1)/2 + (j-i-1).   Absent keys hold RNA_BPP_ABSENT. */
static inline uint64_t rna_bpp_len(uint64_t L) { return L * (L - 1) / 2; }
static inline uint64_t rna_bpp_index(uint64_t L, uint64_t i, uint64_t j) {
  return i * (2 * L - i - 1) / 2 + (j - i - 1);
}

/* ------------------------------------------------------------------------------------------------
 * Batched entry points, HOST buffers (host<->device copies are done inside).
 *
 * Sequences are concatenated base codes (1 byte each, 0..3) with n_seqs+1 offsets.
 * ---------------------------------------------------------------------------------------------- */

/* mccaskill_algo over a batch (+ optionally centroid_fold for n_gammas thresholds, fused on device).
 *   out_logz        [n_seqs]                  sums_external[0][L-1]   (may be NULL)
 *   out_bpp         concatenated packed BPPs, sequence s at bpp_offsets[s] (floats)  (may be NULL).  If the buffer is
 *                   page-locked (cudaHostAlloc / cudaHostRegister) the kernels write it directly while they compute;
 *                   a pageable buffer is filled by a copy after the last kernel.  Same bits either way.
 *   bpp_offsets     [n_seqs+1] or NULL => offsets are the running sum of rna_bpp_len(L_s)
 *   gammas          [n_gammas] centroid thresholds (src/centroid_fold.rs:28)          (n_gammas may be 0)
 *   out_structs     [n_gammas][total_len] dot-bracket bytes '.', '(', ')'; sequence s of gamma g at
 *                   g*total_len + offsets[s]                                         (may be NULL)
 *   out_expect_acc  [n_gammas][n_seqs]  CentroidFold::expect_accuracy                 (may be NULL)
 */
int rna_mccaskill_centroid_batch(rna_handle *h, const uint8_t *bases, const uint32_t *offsets,
                                 uint32_t n_seqs, int model, int allows_short_hairpins,
                                 const float *gammas, uint32_t n_gammas, float *out_logz,
                                 float *out_bpp, const uint64_t *bpp_offsets, uint8_t *out_structs,
                                 float *out_expect_acc);

/* mccaskill_algo only (== rna_mccaskill_centroid_batch with n_gammas = 0). */
int rna_mccaskill_batch(rna_handle *h, const uint8_t *bases, const uint32_t *offsets, uint32_t n_seqs,
                        int model, int allows_short_hairpins, float *out_logz, float *out_bpp,
                        const uint64_t *bpp_offsets);

/* centroid_fold over a batch of packed BPP matrices that already live on the host.
 *   out_pairs       (optional) [n_gammas][total_len] u16 pairs... see rna_centroid_fold for a single one */
int rna_centroid_batch(rna_handle *h, const float *bpp, const uint64_t *bpp_offsets,
                       const uint32_t *offsets, uint32_t n_seqs, const float *gammas, uint32_t n_gammas,
                       uint8_t *out_structs, float *out_expect_acc);

/* durbin_algo over a batch of sequence pairs.  Sequences are given WITHOUT sentinels; the library
 * adds PSEUDO_BASE at both ends exactly like src/bin/durbin_algo.rs:48-50.
 *   pairs           [2*n_pairs] indices (a,b) into the sequence set
 *   out_probs       pair p at prob_offsets[p]: dense row-major (La+2) x (Lb+2) f32, sentinel-indexed,
 *                   zero border (ProbMat of src/durbin_algo.rs:201-242)
 *   prob_offsets    [n_pairs+1] or NULL => running sum of (La+2)*(Lb+2)
 */
int rna_durbin_batch(rna_handle *h, const uint8_t *bases, const uint32_t *offsets, uint32_t n_seqs,
                     const uint32_t *pairs, uint32_t n_pairs, float *out_probs,
                     const uint64_t *prob_offsets);

/* get_fold_sums / get_fold_sums_contra stand-alone (src/mccaskill_algo.rs:282-516), plus the per-pair score memo
 * FoldScores<T> (:13-22, the second value mccaskill_algo returns) — what downstream crates read besides the BPPs.
 * Sequence s receives RNA_SUMS_PLANES planes of rna_sums_len(L_s) floats at sums_offsets[s] (NULL => running sum of
 * RNA_SUMS_PLANES * rna_sums_len(L)); a plane is the upper triangle INCLUDING the diagonal, row-major:
 *   index(i,j) = i*L - i*(i-1)/2 + (j-i),  i <= j.
 * Hash-map members of the reference (sums_close, sums_accessible and the three score maps) hold -inf where the key
 * is absent; dense members hold the reference's values (sums_external is 0.0 on spans the recurrences never visit). */
enum {
  RNA_SUMS_CLOSE = 0,        /* FoldSums::sums_close                                              */
  RNA_SUMS_ACCESSIBLE = 1,   /* FoldSums::sums_accessible                                         */
  RNA_SUMS_EXTERNAL = 2,     /* FoldSums::sums_external                                           */
  RNA_SUMS_RIGHTMOST_EXT = 3,/* FoldSums::sums_rightmost_basepairs_external                       */
  RNA_SUMS_RIGHTMOST_MB = 4, /* FoldSums::sums_rightmost_basepairs_multibranch (CONTRAfold; else -inf) */
  RNA_SUMS_MULTIBRANCH = 5,  /* FoldSums::sums_multibranch                                        */
  RNA_SUMS_1ORMORE = 6,      /* FoldSums::sums_1ormore_basepairs                                  */
  RNA_SCORES_HAIRPIN = 7,    /* FoldScores::hairpin_scores                                        */
  RNA_SCORES_MB_CLOSE = 8,   /* FoldScores::multibranch_close_scores                              */
  RNA_SCORES_ACCESSIBLE = 9, /* FoldScores::accessible_scores                                     */
  RNA_SUMS_PLANES = 10
};
static inline uint64_t rna_sums_len(uint64_t L) { return L * (L + 1) / 2; }
static inline uint64_t rna_sums_index(uint64_t L, uint64_t i, uint64_t j) { return i * L - i * (i - 1) / 2 + (j - i); }
int rna_fold_sums_batch(rna_handle *h, const uint8_t *bases, const uint32_t *offsets, uint32_t n_seqs, int model,
                        int allows_short_hairpins, float *out_sums, const uint64_t *sums_offsets, float *out_logz);

/* FoldScores::twoloop_scores (src/mccaskill_algo.rs:16, filled at :320 / :431): the 4-D memo of two-loop scores, as a
 * flat list in the reference's insertion order — closing pairs (i,j) by span then i ascending, their partners (k,l)
 * with a sums_close entry by k ascending, l descending — value = get_2loop_score(seq, (i,j), (k,l)) bit for bit.
 * One sequence per call (the reference's granularity).  *out_count receives the number of entries; at most `capacity`
 * of them are written (call with capacity 0 to size the buffer: RNA_OK either way).  Reference-exact mode only;
 * lengths up to RNA_MAX_SEQ_LEN (positions are u16 like the reference's HashIndex) and RNA_MAX_FOLD_LEN. */
typedef struct RnaTwoloopScore {
  uint16_t i, j, k, l;
  float score;
} RnaTwoloopScore;
int rna_twoloop_scores(rna_handle *h, const uint8_t *seq, uint32_t seq_len, int model, int allows_short_hairpins,
                       RnaTwoloopScore *out, uint64_t capacity, uint64_t *out_count);

/* ------------------------------------------------------------------------------------------------
 * Single-item convenience wrappers with the reference's per-call granularity.
 * ---------------------------------------------------------------------------------------------- */
int rna_mccaskill_algo(rna_handle *h, const uint8_t *seq, uint32_t seq_len, int uses_contra_model,
                       int allows_short_hairpins, float *out_bpp /* rna_bpp_len(seq_len) */,
                       float *out_logz /* may be NULL */);

/* out_fold_str: seq_len bytes.  out_pairs: up to seq_len/2 (i,j) u16 pairs in the reference's
 * traceback order (basepair_pos_pairs), may be NULL; out_num_pairs may be NULL. */
int rna_centroid_fold(rna_handle *h, const float *bpp, uint32_t seq_len, float centroid_threshold,
                      uint8_t *out_fold_str, uint16_t *out_pairs, uint32_t *out_num_pairs,
                      float *out_expect_accuracy);

int rna_durbin_algo(rna_handle *h, const uint8_t *seq_a, uint32_t len_a, const uint8_t *seq_b,
                    uint32_t len_b, float *out_probs /* (len_a+2)*(len_b+2) */);

/* ------------------------------------------------------------------------------------------------
 * Device-resident variants: every pointer is a DEVICE pointer on the handle's GPU, work is enqueued
 * on `stream` (a cudaStream_t passed as void*; NULL = the handle's own non-blocking stream) and NOT synchronised.
 * Calls of one handle share its scratch memory: each call is ordered behind the previous call of the same handle
 * (an event wait on `stream`), whatever streams they use; rna_set_*_tables waits for the calls in flight.  Inputs
 * must have been validated by the caller (or by rna_validate_bases).  Used for HBM-resident timing
 * and for pipelines that keep the BPPs on the GPU.
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  const uint32_t *h_offsets;      /* HOST copy of the offsets [n_seqs+1]: the launcher buckets by length */
  const uint8_t *d_bases;
  const uint32_t *d_offsets;      /* [n_seqs+1] */
  const uint64_t *d_bpp_offsets;  /* [n_seqs+1], required when d_out_bpp != NULL */
  uint32_t n_seqs;
  uint32_t total_len;
  uint32_t max_len;               /* longest sequence in the batch */
  int model;
  int allows_short_hairpins;
  const float *d_gammas;
  uint32_t n_gammas;
  float *d_out_logz;
  float *d_out_bpp;
  uint8_t *d_out_structs;
  float *d_out_expect_acc;
  uint16_t *d_out_pairs;          /* optional [n_gammas][total_len] (i,j) interleaved u16, traceback order */
  uint32_t *d_out_num_pairs;      /* optional [n_gammas][n_seqs] */
  float *d_out_sums;              /* optional: the planes of rna_fold_sums_batch (reference-exact mode only)   */
  const uint64_t *d_sums_offsets; /* [n_seqs+1], required with d_out_sums                                      */
  int inside_only;                /* stop after the inside pass (no BPP / centroid outputs)                    */
} RnaFoldBatchDev;

int rna_mccaskill_centroid_batch_dev(rna_handle *h, const RnaFoldBatchDev *b, void *stream);

typedef struct {
  const uint32_t *h_offsets;       /* HOST copies: the launcher orders pairs by cost */
  const uint32_t *h_pairs;
  const uint8_t *d_bases;
  const uint32_t *d_offsets;       /* [n_seqs+1] */
  const uint32_t *d_pairs;         /* [2*n_pairs] */
  const uint64_t *d_prob_offsets;  /* [n_pairs+1] */
  uint32_t n_seqs;
  uint32_t n_pairs;
  uint32_t max_len;                /* longest sequence (without sentinels) */
  float *d_out_probs;
} RnaDurbinBatchDev;

int rna_durbin_batch_dev(rna_handle *h, const RnaDurbinBatchDev *b, void *stream);

/* Host-side validation used by all host entry points; exposed for callers of the *_dev variants. */
int rna_validate_bases(const uint8_t *bases, const uint32_t *offsets, uint32_t n_seqs);
/* RNA_ERR_TOO_LONG if a sequence exceeds RNA_MAX_FOLD_LEN (checked by the fold / centroid entry points). */
int rna_validate_fold_lengths(const uint32_t *offsets, uint32_t n_seqs);

/* Length-balanced partition of work units over n_parts GPUs (longest-processing-time-first on the
 * cost model c(L) = L^3 + 500 L^2 for folding, n*m for pairs).  part_of[u] receives
 * the part index of unit u.  Pure host code, no collective. */
int rna_partition_lpt(const uint64_t *costs, uint32_t n_units, uint32_t n_parts, uint32_t *part_of);

/* ------------------------------------------------------------------------------------------------
 * All GPUs of one box behind one object (the reference fans its units out over a thread pool
 * inside the binary: src/bin/centroid_fold.rs:104-161, src/bin/durbin_algo.rs:56-75).  The units of a call are
 * partitioned over the devices by longest-processing-time-first on the same cost model as rna_partition_lpt, one
 * host thread and one rna_handle per device, disjoint output ranges, no collective and no peer traffic.  Same
 * arguments, outputs and status codes as the single-device entry points; results do not depend on the partition.
 *   devices / n_devices : CUDA device ordinals (a device may be listed more than once: that many handles share it);
 *                         n_devices = 0 => every visible device.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rna_multi rna_multi;
int rna_multi_create(const int *devices, int n_devices, rna_multi **out);
int rna_multi_destroy(rna_multi *m);
int rna_multi_num_devices(const rna_multi *m);
rna_handle *rna_multi_handle(rna_multi *m, int k);       /* the k-th per-device handle (stats, last error) */
const char *rna_multi_last_error(const rna_multi *m);
int rna_multi_set_turner_tables(rna_multi *m, const RnaTurnerTables *t);
int rna_multi_set_contra_tables(rna_multi *m, const RnaContraTables *t);
int rna_multi_set_align_tables(rna_multi *m, const RnaAlignTables *t);
int rna_multi_set_numeric_mode(rna_multi *m, int mode);
int rna_multi_mccaskill_centroid_batch(rna_multi *m, const uint8_t *bases, const uint32_t *offsets, uint32_t n_seqs,
                                       int model, int allows_short_hairpins, const float *gammas, uint32_t n_gammas,
                                       float *out_logz, float *out_bpp, const uint64_t *bpp_offsets,
                                       uint8_t *out_structs, float *out_expect_acc);
int rna_multi_durbin_batch(rna_multi *m, const uint8_t *bases, const uint32_t *offsets, uint32_t n_seqs,
                           const uint32_t *pairs, uint32_t n_pairs, float *out_probs, const uint64_t *prob_offsets);
/* Busy time (seconds, host clock around the device's share) of every device in the last rna_multi_* batch call, and
 * the number of units it received: what a caller needs to see the balance of the partition. */
int rna_multi_last_shares(const rna_multi *m, double *busy_seconds /* [n_devices] */, uint32_t *units /* [n_devices] */);

/* ------------------------------------------------------------------------------------------------
 * Thread-SAFE front end of one handle for the reference's call granularity.  The reference's callers fold one
 * sequence per call from the tasks of a thread pool (src/bin/centroid_fold.rs:119-132); one tRNA alone cannot fill a
 * GPU (one CTA, ~3 ms).  Calls that arrive while a launch is in flight are COALESCED: the first caller becomes the
 * leader, gathers every request that is pending with the same (model, allows_short_hairpins, threshold) signature
 * into one rna_mccaskill_centroid_batch call and hands each caller its own results.  Any number of threads may call
 * rna_queue_* on the same queue concurrently; the wrapped handle must not be used directly while the queue is live.
 * ---------------------------------------------------------------------------------------------- */
typedef struct rna_queue rna_queue;
int rna_queue_create(rna_handle *h, rna_queue **out);
int rna_queue_destroy(rna_queue *q);
/* mccaskill_algo (+ centroid_fold when out_fold_str / out_expect_accuracy is given) for ONE sequence; blocks until the
 * caller's results are written.  out_bpp: rna_bpp_len(seq_len) floats or NULL; out_fold_str: seq_len bytes or NULL. */
int rna_queue_mccaskill_algo(rna_queue *q, const uint8_t *seq, uint32_t seq_len, int uses_contra_model,
                             int allows_short_hairpins, float *out_bpp, float *out_logz, float centroid_threshold,
                             uint8_t *out_fold_str, float *out_expect_accuracy);
/* requests served / batched launches issued since the queue was created */
int rna_queue_stats(const rna_queue *q, uint64_t *requests, uint64_t *launches);

/* Counters of the last *_batch call on this handle (kernel launches issued, bytes copied). */
typedef struct {
  uint64_t kernel_launches;
  uint64_t h2d_bytes;
  uint64_t d2h_bytes;
} RnaCallStats;
int rna_get_stats(const rna_handle *h, RnaCallStats *out);

const char *rna_version(void);

/* sizeof() of the blobs as compiled into the library (binding sanity checks). */
size_t rna_sizeof_turner_tables(void);
size_t rna_sizeof_contra_tables(void);
size_t rna_sizeof_align_tables(void);

#ifdef __cplusplus
}
#endif
#endif /* RNA_ALGOS_B200_H */
