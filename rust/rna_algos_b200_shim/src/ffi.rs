//! Hand-written image of include/rna_algos_b200.h (extern "C", plain pointers and sizes).
#![allow(non_camel_case_types, dead_code)]
use std::os::raw::{c_char, c_int, c_void};

pub const RNA_OK: c_int = 0;
pub const RNA_MODEL_TURNER: c_int = 0;
pub const RNA_MODEL_CONTRA: c_int = 1;
pub const RNA_BPP_ABSENT: f32 = -1.0;
pub const RNA_LOOP_TABLE_LEN: usize = 31;
pub const RNA_MAX_SPECIAL_HAIRPINS: usize = 128;
pub const RNA_MAX_SPECIAL_HAIRPIN_LEN: usize = 12;

#[repr(C)]
pub struct rna_handle { _private: [u8; 0] }
#[repr(C)]
pub struct rna_queue { _private: [u8; 0] }
#[repr(C)]
pub struct rna_multi { _private: [u8; 0] }
pub const RNA_SUMS_PLANES: usize = 10;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct RnaSpecialHairpin {
    pub len: u8,
    pub seq: [u8; RNA_MAX_SPECIAL_HAIRPIN_LEN],
    pub _pad: [u8; 3],
    pub score: f32,
}

type T4 = [[[[f32; 4]; 4]; 4]; 4];
type T3 = [[[f32; 4]; 4]; 4];

/// rna_ss_params::compiled_scores_turner + the model caps (field names = the symbols src/utils.rs consumes).
#[repr(C)]
pub struct RnaTurnerTables {
    pub max_2loop_len: i32,
    pub min_span_hairpin_close: i32,
    pub min_hairpin_len: i32,
    pub max_hairpin_len_extrapolation: i32,
    pub min_hairpin_len_extrapolation: i32,
    pub num_special_hairpins: i32,
    pub coeff_hairpin_len_extrapolation: f32,
    pub helix_augu_end_penalty: f32,
    pub ninio_coeff: f32,
    pub ninio_max: f32,
    pub init_multibranch_base: f32,
    pub coeff_num_branches: f32,
    pub hairpin_scores_init: [f32; RNA_LOOP_TABLE_LEN],
    pub bulge_scores_init: [f32; RNA_LOOP_TABLE_LEN],
    pub interior_scores_init: [f32; RNA_LOOP_TABLE_LEN],
    pub stack_scores: T4,
    pub terminal_mismatch_scores_hairpin: T4,
    pub terminal_mismatch_scores_1xmany: T4,
    pub terminal_mismatch_scores_2x3: T4,
    pub terminal_mismatch_scores_interior: T4,
    pub terminal_mismatch_scores_multibranch: T4,
    pub dangling_scores_5prime: T3,
    pub dangling_scores_3prime: T3,
    pub interior_scores_1x1: [[T4; 4]; 4],
    pub interior_scores_1x2: [[[T4; 4]; 4]; 4],
    pub interior_scores_2x2: [[[[T4; 4]; 4]; 4]; 4],
    pub hairpin_scores_special: [RnaSpecialHairpin; RNA_MAX_SPECIAL_HAIRPINS],
}

/// Field-for-field image of FoldScoreSets (src/utils.rs:91-119) plus the caps the recurrences read.
#[repr(C)]
pub struct RnaContraTables {
    pub max_loop_len: i32,
    pub min_span_hairpin_close: i32,
    pub max_interior_explicit: i32,
    pub _pad: i32,
    pub hairpin_scores_len: [f32; 31],
    pub bulge_scores_len: [f32; 30],
    pub interior_scores_len: [f32; 29],
    pub interior_scores_symmetric: [f32; 15],
    pub interior_scores_asymmetric: [f32; 28],
    pub stack_scores: T4,
    pub terminal_mismatch_scores: T4,
    pub dangling_scores_left: T3,
    pub dangling_scores_right: T3,
    pub helix_close_scores: [[f32; 4]; 4],
    pub basepair_scores: [[f32; 4]; 4],
    pub interior_scores_explicit: [[f32; 4]; 4],
    pub bulge_scores_0x1: [f32; 4],
    pub interior_scores_1x1: [[f32; 4]; 4],
    pub multibranch_score_base: f32,
    pub multibranch_score_basepair: f32,
    pub multibranch_score_unpair: f32,
    pub external_score_basepair: f32,
    pub external_score_unpair: f32,
    pub hairpin_scores_len_cumulative: [f32; 31],
    pub bulge_scores_len_cumulative: [f32; 30],
    pub interior_scores_len_cumulative: [f32; 29],
    pub interior_scores_symmetric_cumulative: [f32; 15],
    pub interior_scores_asymmetric_cumulative: [f32; 28],
}

/// Image of AlignScores (src/durbin_algo.rs:4-14).
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RnaAlignTables {
    pub match2match_score: f32,
    pub match2insert_score: f32,
    pub insert_extend_score: f32,
    pub insert_switch_score: f32,
    pub init_match_score: f32,
    pub init_insert_score: f32,
    pub insert_scores: [f32; 4],
    pub match_scores: [[f32; 4]; 4],
}

/// include/rna_algos_b200.h RnaTwoloopScore: one entry of FoldScores::twoloop_scores
#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct RnaTwoloopScore { pub i: u16, pub j: u16, pub k: u16, pub l: u16, pub score: f32 }

extern "C" {
    pub fn rna_create(device: c_int, out: *mut *mut rna_handle) -> c_int;
    pub fn rna_destroy(h: *mut rna_handle) -> c_int;
    pub fn rna_last_error(h: *const rna_handle) -> *const c_char;
    pub fn rna_set_turner_tables(h: *mut rna_handle, t: *const RnaTurnerTables) -> c_int;
    pub fn rna_set_contra_tables(h: *mut rna_handle, t: *const RnaContraTables) -> c_int;
    pub fn rna_set_align_tables(h: *mut rna_handle, t: *const RnaAlignTables) -> c_int;
    pub fn rna_contra_tables_accumulate(t: *mut RnaContraTables);
    pub fn rna_align_tables_contralign_v201(t: *mut RnaAlignTables);
    pub fn rna_sizeof_turner_tables() -> usize;
    pub fn rna_sizeof_contra_tables() -> usize;
    pub fn rna_sizeof_align_tables() -> usize;

    pub fn rna_mccaskill_centroid_batch(h: *mut rna_handle, bases: *const u8, offsets: *const u32, n_seqs: u32,
        model: c_int, allows_short_hairpins: c_int, gammas: *const f32, n_gammas: u32, out_logz: *mut f32,
        out_bpp: *mut f32, bpp_offsets: *const u64, out_structs: *mut u8, out_expect_acc: *mut f32) -> c_int;
    pub fn rna_mccaskill_batch(h: *mut rna_handle, bases: *const u8, offsets: *const u32, n_seqs: u32, model: c_int,
        allows_short_hairpins: c_int, out_logz: *mut f32, out_bpp: *mut f32, bpp_offsets: *const u64) -> c_int;
    pub fn rna_centroid_batch(h: *mut rna_handle, bpp: *const f32, bpp_offsets: *const u64, offsets: *const u32,
        n_seqs: u32, gammas: *const f32, n_gammas: u32, out_structs: *mut u8, out_expect_acc: *mut f32) -> c_int;
    pub fn rna_durbin_batch(h: *mut rna_handle, bases: *const u8, offsets: *const u32, n_seqs: u32, pairs: *const u32,
        n_pairs: u32, out_probs: *mut f32, prob_offsets: *const u64) -> c_int;

    pub fn rna_mccaskill_algo(h: *mut rna_handle, seq: *const u8, seq_len: u32, uses_contra_model: c_int,
        allows_short_hairpins: c_int, out_bpp: *mut f32, out_logz: *mut f32) -> c_int;
    pub fn rna_centroid_fold(h: *mut rna_handle, bpp: *const f32, seq_len: u32, centroid_threshold: f32,
        out_fold_str: *mut u8, out_pairs: *mut u16, out_num_pairs: *mut u32, out_expect_accuracy: *mut f32) -> c_int;
    pub fn rna_durbin_algo(h: *mut rna_handle, seq_a: *const u8, len_a: u32, seq_b: *const u8, len_b: u32,
        out_probs: *mut f32) -> c_int;
    pub fn rna_fold_sums_batch(h: *mut rna_handle, bases: *const u8, offsets: *const u32, n_seqs: u32, model: c_int,
        allows_short_hairpins: c_int, out_sums: *mut f32, sums_offsets: *const u64, out_logz: *mut f32) -> c_int;
    pub fn rna_twoloop_scores(h: *mut rna_handle, seq: *const u8, seq_len: u32, model: c_int, allows_short_hairpins: c_int,
        out: *mut RnaTwoloopScore, capacity: u64, out_count: *mut u64) -> c_int;
    pub fn rna_set_numeric_mode(h: *mut rna_handle, mode: c_int) -> c_int;
    pub fn rna_queue_create(h: *mut rna_handle, out: *mut *mut rna_queue) -> c_int;
    pub fn rna_queue_destroy(q: *mut rna_queue) -> c_int;
    pub fn rna_queue_mccaskill_algo(q: *mut rna_queue, seq: *const u8, seq_len: u32, uses_contra_model: c_int,
        allows_short_hairpins: c_int, out_bpp: *mut f32, out_logz: *mut f32, centroid_threshold: f32,
        out_fold_str: *mut u8, out_expect_accuracy: *mut f32) -> c_int;
    pub fn rna_multi_create(devices: *const c_int, n_devices: c_int, out: *mut *mut rna_multi) -> c_int;
    pub fn rna_multi_destroy(m: *mut rna_multi) -> c_int;
    pub fn rna_multi_set_turner_tables(m: *mut rna_multi, t: *const RnaTurnerTables) -> c_int;
    pub fn rna_multi_set_contra_tables(m: *mut rna_multi, t: *const RnaContraTables) -> c_int;
    pub fn rna_multi_set_align_tables(m: *mut rna_multi, t: *const RnaAlignTables) -> c_int;
    pub fn rna_multi_mccaskill_centroid_batch(m: *mut rna_multi, bases: *const u8, offsets: *const u32, n_seqs: u32,
        model: c_int, allows_short_hairpins: c_int, gammas: *const f32, n_gammas: u32, out_logz: *mut f32,
        out_bpp: *mut f32, bpp_offsets: *const u64, out_structs: *mut u8, out_expect_acc: *mut f32) -> c_int;
    pub fn rna_multi_durbin_batch(m: *mut rna_multi, bases: *const u8, offsets: *const u32, n_seqs: u32, pairs: *const u32,
        n_pairs: u32, out_probs: *mut f32, prob_offsets: *const u64) -> c_int;
    pub fn rna_validate_bases(bases: *const u8, offsets: *const u32, n_seqs: u32) -> c_int;
    pub fn rna_partition_lpt(costs: *const u64, n_units: u32, n_parts: u32, part_of: *mut u32) -> c_int;
}

#[allow(unused)]
fn _unused(_: *mut c_void) {}
