//! Reference-shaped API over the CUDA library: the three public functions of heartsh/rna-algos
//! (`mccaskill_algo`, `centroid_fold`, `durbin_algo`; src/mccaskill_algo.rs:247, src/centroid_fold.rs:25,
//! src/durbin_algo.rs:73) with the same argument meaning and output types, plus batched calls.
//!
//! UNTESTED: the repository's build image has no Rust toolchain.  The C ABI underneath is tested
//! (tests/, through ctypes and through the C++ command-line front ends).
//!
//! There is no CPU path: `Gpu::new` fails when no CUDA device is usable.
pub mod ffi;
pub mod reference_api;   // the reference's own free-function signatures (per-call score sets, coalesced single calls)

use std::collections::HashMap;
use std::convert::TryFrom;
use std::ffi::CStr;
use std::hash::Hash;

pub type Prob = f32;
pub type Base = usize;                               // 0..3 = A C G U, 4 = PSEUDO_BASE (src/utils.rs:84,122)
pub type SparseProbMat<T> = HashMap<(T, T), Prob>;   // hashbrown in the reference; same keys and values
pub type ProbMat = Vec<Vec<Prob>>;

/// CentroidFold<T> of src/centroid_fold.rs:3-23.
pub struct CentroidFold<T> {
    pub basepair_pos_pairs: Vec<(T, T)>,
    pub expect_accuracy: Prob,
}

/// The index types the reference instantiates (u8 / u16: src/bin/centroid_fold.rs:85-101).
pub trait HashIndex: Copy + Eq + Hash + TryFrom<usize> + Into<usize> {}
impl<T: Copy + Eq + Hash + TryFrom<usize> + Into<usize>> HashIndex for T {}

#[derive(Debug)]
pub struct RnaError { pub code: i32, pub detail: String }

pub struct Gpu { h: *mut ffi::rna_handle }
unsafe impl Send for Gpu {}

impl Drop for Gpu {
    fn drop(&mut self) { unsafe { ffi::rna_destroy(self.h); } }
}

impl Gpu {
    /// One handle per device; one in-flight call per handle.
    pub fn new(device: i32) -> Result<Gpu, RnaError> {
        let mut h = std::ptr::null_mut();
        let rc = unsafe { ffi::rna_create(device, &mut h) };
        if rc != ffi::RNA_OK { return Err(RnaError { code: rc, detail: "no usable CUDA device (there is no CPU path)".into() }); }
        assert_eq!(unsafe { ffi::rna_sizeof_turner_tables() }, std::mem::size_of::<ffi::RnaTurnerTables>());
        assert_eq!(unsafe { ffi::rna_sizeof_contra_tables() }, std::mem::size_of::<ffi::RnaContraTables>());
        assert_eq!(unsafe { ffi::rna_sizeof_align_tables() }, std::mem::size_of::<ffi::RnaAlignTables>());
        let g = Gpu { h };
        let mut a = ffi::RnaAlignTables::default();
        unsafe { ffi::rna_align_tables_contralign_v201(&mut a) };   // src/compiled_align_scores.rs:2-19
        g.check(unsafe { ffi::rna_set_align_tables(g.h, &a) })?;
        Ok(g)
    }
    fn check(&self, rc: i32) -> Result<(), RnaError> {
        if rc == ffi::RNA_OK { return Ok(()); }
        let detail = unsafe { CStr::from_ptr(ffi::rna_last_error(self.h)) }.to_string_lossy().into_owned();
        Err(RnaError { code: rc, detail })
    }
    /// Copy the rna-ss-params constants into the blobs field by field (names are identical) and hand them over;
    /// `contra` must hold the *_len ("at least") arrays, the cumulative ones are filled here
    /// (FoldScoreSets::accumulate, src/mccaskill_algo.rs:60-86).
    pub fn set_tables(&self, turner: &ffi::RnaTurnerTables, contra: &mut ffi::RnaContraTables) -> Result<(), RnaError> {
        unsafe { ffi::rna_contra_tables_accumulate(contra) };
        self.check(unsafe { ffi::rna_set_turner_tables(self.h, turner) })?;
        self.check(unsafe { ffi::rna_set_contra_tables(self.h, contra) })
    }

    /// mccaskill_algo<T> (src/mccaskill_algo.rs:247-255).  The FoldScores second element is not produced: no
    /// in-tree caller reads it (src/bin/mccaskill_algo.rs:78, src/bin/centroid_fold.rs:129, tests/tests.rs:31).
    pub fn mccaskill_algo<T: HashIndex>(&self, seq: &[Base], uses_contra_model: bool, allows_short_hairpins: bool)
        -> Result<SparseProbMat<T>, RnaError> {
        let l = seq.len();
        let bases: Vec<u8> = seq.iter().map(|&b| b as u8).collect();
        let mut bpp = vec![0f32; l * l.saturating_sub(1) / 2];
        self.check(unsafe { ffi::rna_mccaskill_algo(self.h, bases.as_ptr(), l as u32, uses_contra_model as i32,
            allows_short_hairpins as i32, bpp.as_mut_ptr(), std::ptr::null_mut()) })?;
        let mut m = SparseProbMat::<T>::default();
        for i in 0..l { for j in i + 1..l {
            let p = bpp[i * (2 * l - i - 1) / 2 + (j - i - 1)];                 // rna_bpp_index
            if p != ffi::RNA_BPP_ABSENT {                                       // keys present with 0.0 are kept
                if let (Ok(a), Ok(b)) = (T::try_from(i), T::try_from(j)) { m.insert((a, b), p); }
            }
        }}
        Ok(m)
    }

    /// centroid_fold<T> (src/centroid_fold.rs:25-32); pairs come back in the reference's traceback order.
    pub fn centroid_fold<T: HashIndex>(&self, basepair_probs: &SparseProbMat<T>, seq_len: usize, centroid_threshold: Prob)
        -> Result<CentroidFold<T>, RnaError> {
        let mut bpp = vec![ffi::RNA_BPP_ABSENT; seq_len * seq_len.saturating_sub(1) / 2];
        for (&(i, j), &p) in basepair_probs {
            let (i, j): (usize, usize) = (i.into(), j.into());
            bpp[i * (2 * seq_len - i - 1) / 2 + (j - i - 1)] = p;
        }
        let mut s = vec![0u8; seq_len];
        let mut pairs = vec![0u16; 2 * seq_len];
        let (mut n, mut ea) = (0u32, 0f32);
        self.check(unsafe { ffi::rna_centroid_fold(self.h, bpp.as_ptr(), seq_len as u32, centroid_threshold,
            s.as_mut_ptr(), pairs.as_mut_ptr(), &mut n, &mut ea) })?;
        let mut f = CentroidFold { basepair_pos_pairs: Vec::with_capacity(n as usize), expect_accuracy: ea };
        for k in 0..n as usize {
            if let (Ok(a), Ok(b)) = (T::try_from(pairs[2 * k] as usize), T::try_from(pairs[2 * k + 1] as usize)) {
                f.basepair_pos_pairs.push((a, b));
            }
        }
        Ok(f)
    }

    /// durbin_algo (src/durbin_algo.rs:73).  Callers pass PSEUDO_BASE-padded sequences
    /// (src/bin/durbin_algo.rs:48-50); the library adds the sentinels itself, so they are stripped here.
    pub fn durbin_algo(&self, seq_pair: &(&[Base], &[Base])) -> Result<ProbMat, RnaError> {
        let strip = |s: &[Base]| s[1..s.len() - 1].iter().map(|&b| b as u8).collect::<Vec<u8>>();
        let (a, b) = (strip(seq_pair.0), strip(seq_pair.1));
        let (n, m) = (a.len() + 2, b.len() + 2);
        let mut flat = vec![0f32; n * m];
        self.check(unsafe { ffi::rna_durbin_algo(self.h, a.as_ptr(), a.len() as u32, b.as_ptr(), b.len() as u32,
            flat.as_mut_ptr()) })?;
        Ok(flat.chunks(m).map(|r| r.to_vec()).collect())                        // dense n x m, zero border
    }

    /// What src/bin/centroid_fold.rs:104-161 does with a thread pool, as ONE call: BPP + dot-bracket strings for all
    /// sequences and all thresholds.  Returns structs[g][s] as bytes '.', '(', ')'.
    pub fn mccaskill_centroid_batch(&self, seqs: &[Vec<Base>], uses_contra_model: bool, gammas: &[Prob])
        -> Result<Vec<Vec<Vec<u8>>>, RnaError> {
        let mut bases = Vec::new();
        let mut offsets = vec![0u32];
        for s in seqs { bases.extend(s.iter().map(|&b| b as u8)); offsets.push(bases.len() as u32); }
        let total = bases.len();
        let mut structs = vec![0u8; gammas.len() * total];
        self.check(unsafe { ffi::rna_mccaskill_centroid_batch(self.h, bases.as_ptr(), offsets.as_ptr(), seqs.len() as u32,
            if uses_contra_model { ffi::RNA_MODEL_CONTRA } else { ffi::RNA_MODEL_TURNER }, 0, gammas.as_ptr(),
            gammas.len() as u32, std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null(), structs.as_mut_ptr(),
            std::ptr::null_mut()) })?;
        Ok((0..gammas.len()).map(|g| (0..seqs.len()).map(|s|
            structs[g * total + offsets[s] as usize..g * total + offsets[s + 1] as usize].to_vec()).collect()).collect())
    }
}
