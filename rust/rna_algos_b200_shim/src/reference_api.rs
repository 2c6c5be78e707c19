//! The reference crate's three entry points with the reference's OWN signatures
//! (src/mccaskill_algo.rs:247-255, src/centroid_fold.rs:25-32, src/durbin_algo.rs:73), so that a downstream crate
//! switches by changing a `use` line:
//!
//!   pub fn mccaskill_algo<T>(seq, uses_contra_model, allows_short_hairpins, fold_score_sets: &FoldScoreSets)
//!       -> (SparseProbMat<T>, FoldScores<T>)
//!   pub fn centroid_fold<T>(basepair_probs: &SparseProbMat<T>, seq_len, centroid_threshold) -> CentroidFold<T>
//!   pub fn durbin_algo(seq_pair: &SeqPair, align_scores: &AlignScores) -> ProbMat
//!
//! The score sets are PER-CALL arguments like in the reference: each distinct table blob (keyed by a hash of its
//! bytes) gets a GPU handle of its own with those tables uploaded once, so a caller that trains or perturbs scores
//! gets what it passed, never another caller's tables.  Single calls from the tasks of a thread pool
//! (src/bin/centroid_fold.rs:119-132) go through `rna_queue`, which coalesces them into batched launches.
//! Like the reference, these functions panic on invalid input (non-ACGU base, empty sequence): src/utils.rs:570-572.
//!
//! UNTESTED: no Rust toolchain in the repository's build image.
use crate::ffi;
use crate::{Base, CentroidFold, HashIndex, Prob, ProbMat, SparseProbMat};
use std::collections::hash_map::DefaultHasher;
use std::collections::HashMap;
use std::hash::{Hash, Hasher};
use std::sync::Mutex;

pub type Score = f32;
pub type SeqSlice<'a> = &'a [Base];
pub type SeqPair<'a> = (SeqSlice<'a>, SeqSlice<'a>);
pub type SparseScoreMat<T> = HashMap<(T, T), Score>;
pub type ScoreMat4d<T> = HashMap<(T, T, T, T), Score>;

/// src/utils.rs:91-119, field for field (= the layout of ffi::RnaContraTables after the four caps).
pub type FoldScoreSets = ffi::RnaContraTables;
/// src/durbin_algo.rs:4-14
pub type AlignScores = ffi::RnaAlignTables;

/// src/mccaskill_algo.rs:13-22: all four members are filled (`twoloop_scores` through rna_twoloop_scores).
pub struct FoldScores<T: Hash + Eq> {
    pub hairpin_scores: SparseScoreMat<T>,
    pub twoloop_scores: ScoreMat4d<T>,
    pub multibranch_close_scores: SparseScoreMat<T>,
    pub accessible_scores: SparseScoreMat<T>,
}

struct Slot { h: *mut ffi::rna_handle, q: *mut ffi::rna_queue }
unsafe impl Send for Slot {}

fn bytes_of<S>(s: &S) -> &[u8] { unsafe { std::slice::from_raw_parts(s as *const S as *const u8, std::mem::size_of::<S>()) } }
fn key_of<S>(s: &S) -> u64 { let mut h = DefaultHasher::new(); bytes_of(s).hash(&mut h); h.finish() }

static SLOTS: Mutex<Option<HashMap<u64, Slot>>> = Mutex::new(None);
/// The genuine Turner 2004 constants are compile-time consts of rna-ss-params in the reference; a shim built against
/// that crate fills this blob once (tools/ref_dump shows the field mapping) before the first Turner call.
pub static TURNER: Mutex<Option<Box<ffi::RnaTurnerTables>>> = Mutex::new(None);

fn slot_for(key: u64, setup: impl FnOnce(*mut ffi::rna_handle)) -> (*mut ffi::rna_handle, *mut ffi::rna_queue) {
    let mut g = SLOTS.lock().unwrap();
    let map = g.get_or_insert_with(HashMap::new);
    if let Some(s) = map.get(&key) { return (s.h, s.q); }
    let mut h = std::ptr::null_mut();
    assert_eq!(unsafe { ffi::rna_create(0, &mut h) }, ffi::RNA_OK, "no usable CUDA device (there is no CPU path)");
    setup(h);
    let mut q = std::ptr::null_mut();
    assert_eq!(unsafe { ffi::rna_queue_create(h, &mut q) }, ffi::RNA_OK);
    map.insert(key, Slot { h, q });
    (h, q)
}

pub fn mccaskill_algo<T: HashIndex>(seq: SeqSlice, uses_contra_model: bool, allows_short_hairpins: bool,
                                    fold_score_sets: &FoldScoreSets) -> (SparseProbMat<T>, FoldScores<T>) {
    let (h, q) = slot_for(key_of(fold_score_sets), |h| unsafe {
        assert_eq!(ffi::rna_set_contra_tables(h, fold_score_sets), ffi::RNA_OK);
        if let Some(t) = TURNER.lock().unwrap().as_ref() { assert_eq!(ffi::rna_set_turner_tables(h, &**t), ffi::RNA_OK); }
    });
    let l = seq.len();
    let bases: Vec<u8> = seq.iter().map(|&b| u8::try_from(b).expect("base code")).collect();
    let mut bpp = vec![0f32; l * l.saturating_sub(1) / 2];
    let rc = unsafe { ffi::rna_queue_mccaskill_algo(q, bases.as_ptr(), l as u32, uses_contra_model as i32,
        allows_short_hairpins as i32, bpp.as_mut_ptr(), std::ptr::null_mut(), 0.0, std::ptr::null_mut(), std::ptr::null_mut()) };
    assert_eq!(rc, ffi::RNA_OK, "rna_queue_mccaskill_algo failed");
    let mut probs = SparseProbMat::<T>::default();
    for i in 0..l { for j in i + 1..l {
        let p = bpp[i * (2 * l - i - 1) / 2 + (j - i - 1)];
        if p != ffi::RNA_BPP_ABSENT { if let (Ok(a), Ok(b)) = (T::try_from(i), T::try_from(j)) { probs.insert((a, b), p); } }
    }}
    // FoldScores: planes 7..9 of rna_fold_sums_batch (hash-map members: -inf = key absent)
    let pl = l * (l + 1) / 2;
    let mut planes = vec![0f32; ffi::RNA_SUMS_PLANES * pl];
    let off = [0u32, l as u32];
    let rc = unsafe { ffi::rna_fold_sums_batch(h, bases.as_ptr(), off.as_ptr(), 1,
        if uses_contra_model { ffi::RNA_MODEL_CONTRA } else { ffi::RNA_MODEL_TURNER }, allows_short_hairpins as i32,
        planes.as_mut_ptr(), std::ptr::null(), std::ptr::null_mut()) };
    assert_eq!(rc, ffi::RNA_OK, "rna_fold_sums_batch failed");
    let mut fs = FoldScores::<T> { hairpin_scores: HashMap::new(), twoloop_scores: HashMap::new(),
                                   multibranch_close_scores: HashMap::new(), accessible_scores: HashMap::new() };
    for i in 0..l { for j in i..l {
        let x = i * l - i * i.saturating_sub(1) / 2 + (j - i);
        if let (Ok(a), Ok(b)) = (T::try_from(i), T::try_from(j)) {
            if planes[7 * pl + x] > f32::NEG_INFINITY { fs.hairpin_scores.insert((a, b), planes[7 * pl + x]); }
            if planes[8 * pl + x] > f32::NEG_INFINITY { fs.multibranch_close_scores.insert((a, b), planes[8 * pl + x]); }
            if planes[9 * pl + x] > f32::NEG_INFINITY { fs.accessible_scores.insert((a, b), planes[9 * pl + x]); }
        }
    }}
    // twoloop_scores: the 4-D memo (src/mccaskill_algo.rs:320, 431), sized by a first call with capacity 0
    let model = if uses_contra_model { ffi::RNA_MODEL_CONTRA } else { ffi::RNA_MODEL_TURNER };
    let mut cnt = 0u64;
    let rc = unsafe { ffi::rna_twoloop_scores(h, bases.as_ptr(), l as u32, model, allows_short_hairpins as i32,
        std::ptr::null_mut(), 0, &mut cnt) };
    assert_eq!(rc, ffi::RNA_OK, "rna_twoloop_scores failed");
    let mut ent = vec![ffi::RnaTwoloopScore::default(); cnt as usize];
    if cnt > 0 {
        let rc = unsafe { ffi::rna_twoloop_scores(h, bases.as_ptr(), l as u32, model, allows_short_hairpins as i32,
            ent.as_mut_ptr(), cnt, &mut cnt) };
        assert_eq!(rc, ffi::RNA_OK, "rna_twoloop_scores failed");
    }
    for e in &ent {
        if let (Ok(a), Ok(b), Ok(c), Ok(d)) = (T::try_from(e.i as usize), T::try_from(e.j as usize), T::try_from(e.k as usize), T::try_from(e.l as usize)) {
            fs.twoloop_scores.insert((a, b, c, d), e.score);
        }
    }
    (probs, fs)
}

pub fn centroid_fold<T: HashIndex>(basepair_probs: &SparseProbMat<T>, seq_len: usize, centroid_threshold: Prob) -> CentroidFold<T> {
    // (no score tables involved: any handle will do; key 0 = the shared default)
    let (h, _) = slot_for(0, |_| {});
    let mut bpp = vec![ffi::RNA_BPP_ABSENT; seq_len * seq_len.saturating_sub(1) / 2];
    for (&(i, j), &p) in basepair_probs {
        let (i, j): (usize, usize) = (i.into(), j.into());
        bpp[i * (2 * seq_len - i - 1) / 2 + (j - i - 1)] = p;
    }
    let mut s = vec![0u8; seq_len];
    let mut pairs = vec![0u16; 2 * seq_len];
    let (mut n, mut ea) = (0u32, 0f32);
    let _g = SLOTS.lock().unwrap();   // (a bare handle is one call at a time)
    let rc = unsafe { ffi::rna_centroid_fold(h, bpp.as_ptr(), seq_len as u32, centroid_threshold, s.as_mut_ptr(),
        pairs.as_mut_ptr(), &mut n, &mut ea) };
    assert_eq!(rc, ffi::RNA_OK, "rna_centroid_fold failed");
    let mut f = CentroidFold { basepair_pos_pairs: Vec::with_capacity(n as usize), expect_accuracy: ea };
    for k in 0..n as usize {
        if let (Ok(a), Ok(b)) = (T::try_from(pairs[2 * k] as usize), T::try_from(pairs[2 * k + 1] as usize)) { f.basepair_pos_pairs.push((a, b)); }
    }
    f
}

pub fn durbin_algo(seq_pair: &SeqPair, align_scores: &AlignScores) -> ProbMat {
    let (h, _) = slot_for(key_of(align_scores) ^ 0x9e3779b97f4a7c15, |h| unsafe {
        assert_eq!(ffi::rna_set_align_tables(h, align_scores), ffi::RNA_OK);
    });
    // callers pass PSEUDO_BASE-padded sequences (src/bin/durbin_algo.rs:48-50); the library adds the sentinels itself
    let strip = |s: &[Base]| s[1..s.len() - 1].iter().map(|&b| b as u8).collect::<Vec<u8>>();
    let (a, b) = (strip(seq_pair.0), strip(seq_pair.1));
    let (n, m) = (a.len() + 2, b.len() + 2);
    let mut flat = vec![0f32; n * m];
    let _g = SLOTS.lock().unwrap();
    let rc = unsafe { ffi::rna_durbin_algo(h, a.as_ptr(), a.len() as u32, b.as_ptr(), b.len() as u32, flat.as_mut_ptr()) };
    assert_eq!(rc, ffi::RNA_OK, "rna_durbin_algo failed");
    flat.chunks(m).map(|r| r.to_vec()).collect()
}
