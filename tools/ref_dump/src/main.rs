// ref_dump — see Cargo.toml.  UNTESTED IN THIS REPOSITORY'S IMAGE (no rustc); written against the public items of
// heartsh/rna-algos 0.1.37 that the hot path uses (src/mccaskill_algo.rs, src/centroid_fold.rs, src/durbin_algo.rs,
// src/utils.rs) and the rna-ss-params symbols those files import.  Where the element type of an upstream table is
// not visible from the reference tree, values go through `as f32` / `as u8`.
extern crate bio;
extern crate rna_algos;

use rna_algos::centroid_fold::*;
use rna_algos::durbin_algo::*;
use rna_algos::mccaskill_algo::*;
use rna_algos::utils::*;
use std::fs::{create_dir_all, File};
use std::io::Write;
use std::path::Path;

fn put_i32(b: &mut Vec<u8>, x: i32) { b.extend_from_slice(&x.to_le_bytes()); }
fn put_f32(b: &mut Vec<u8>, x: f32) { b.extend_from_slice(&x.to_le_bytes()); }
fn put_pad(b: &mut Vec<u8>, n_floats: usize, have: usize) { for _ in have..n_floats { put_f32(b, 0.); } }

// "RNATBL01" | u32 kind | u32 size | struct bytes   (rna_algos_b200/tables.py: dump_table_file)
fn write_blob(path: &Path, kind: u32, body: &[u8]) {
  let mut f = File::create(path).unwrap();
  f.write_all(b"RNATBL01").unwrap();
  f.write_all(&kind.to_le_bytes()).unwrap();
  f.write_all(&(body.len() as u32).to_le_bytes()).unwrap();
  f.write_all(body).unwrap();
}

// NPY 1.0, little-endian, C order
fn write_npy(path: &Path, descr: &str, shape: &[usize], data: &[u8]) {
  let dims = shape.iter().map(|d| format!("{},", d)).collect::<Vec<_>>().join(" ");
  let mut h = format!("{{'descr': '{}', 'fortran_order': False, 'shape': ({}), }}", descr, dims);
  while (10 + h.len() + 1) % 64 != 0 { h.push(' '); }
  h.push('\n');
  let mut f = File::create(path).unwrap();
  f.write_all(b"\x93NUMPY\x01\x00").unwrap();
  f.write_all(&(h.len() as u16).to_le_bytes()).unwrap();
  f.write_all(h.as_bytes()).unwrap();
  f.write_all(data).unwrap();
}
fn f32s(v: &[f32]) -> Vec<u8> { v.iter().flat_map(|x| x.to_le_bytes().to_vec()).collect() }

// RnaTurnerTables, field for field (include/rna_algos_b200.h)
fn turner_blob() -> Vec<u8> {
  let mut b = Vec::new();
  put_i32(&mut b, MAX_2LOOP_LEN as i32);
  put_i32(&mut b, MIN_SPAN_HAIRPIN_CLOSE as i32);
  put_i32(&mut b, MIN_HAIRPIN_LEN as i32);
  put_i32(&mut b, MAX_HAIRPIN_LEN_EXTRAPOLATION as i32);
  put_i32(&mut b, MIN_HAIRPIN_LEN_EXTRAPOLATION as i32);
  put_i32(&mut b, HAIRPIN_SCORES_SPECIAL.len() as i32);
  put_f32(&mut b, COEFF_HAIRPIN_LEN_EXTRAPOLATION as f32);
  put_f32(&mut b, HELIX_AUGU_END_PENALTY as f32);
  put_f32(&mut b, NINIO_COEFF as f32);
  put_f32(&mut b, NINIO_MAX as f32);
  put_f32(&mut b, INIT_MULTIBRANCH_BASE as f32);
  put_f32(&mut b, COEFF_NUM_BRANCHES as f32);
  for t in [&HAIRPIN_SCORES_INIT[..], &BULGE_SCORES_INIT[..], &INTERIOR_SCORES_INIT[..]].iter() {
    for x in t.iter().take(31) { put_f32(&mut b, *x as f32); }
    put_pad(&mut b, 31, t.len().min(31));
  }
  for t in [&STACK_SCORES, &TERMINAL_MISMATCH_SCORES_HAIRPIN, &TERMINAL_MISMATCH_SCORES_1XMANY,
            &TERMINAL_MISMATCH_SCORES_2X3, &TERMINAL_MISMATCH_SCORES_INTERIOR, &TERMINAL_MISMATCH_SCORES_MULTIBRANCH].iter() {
    for w in t.iter() { for x in w.iter() { for y in x.iter() { for z in y.iter() { put_f32(&mut b, *z as f32); } } } }
  }
  for t in [&DANGLING_SCORES_5PRIME, &DANGLING_SCORES_3PRIME].iter() {
    for w in t.iter() { for x in w.iter() { for y in x.iter() { put_f32(&mut b, *y as f32); } } }
  }
  for a in INTERIOR_SCORES_1X1.iter() { for c in a.iter() { for d in c.iter() { for e in d.iter() { for f in e.iter() { for g in f.iter() {
    put_f32(&mut b, *g as f32); } } } } } }
  for a in INTERIOR_SCORES_1X2.iter() { for c in a.iter() { for d in c.iter() { for e in d.iter() { for f in e.iter() { for g in f.iter() { for h in g.iter() {
    put_f32(&mut b, *h as f32); } } } } } } }
  for a in INTERIOR_SCORES_2X2.iter() { for c in a.iter() { for d in c.iter() { for e in d.iter() { for f in e.iter() { for g in f.iter() { for h in g.iter() { for i in h.iter() {
    put_f32(&mut b, *i as f32); } } } } } } } }
  // RnaSpecialHairpin[128]: u8 len | u8 seq[12] | u8 pad[3] | f32 score
  assert!(HAIRPIN_SCORES_SPECIAL.len() <= 128);
  for n in 0..128 {
    if n < HAIRPIN_SCORES_SPECIAL.len() {
      let e = &HAIRPIN_SCORES_SPECIAL[n];
      assert!(e.0.len() <= 12);
      b.push(e.0.len() as u8);
      for k in 0..12 { b.push(if k < e.0.len() { e.0[k] as u8 } else { 0 }); }
      b.extend_from_slice(&[0u8; 3]);
      put_f32(&mut b, e.1 as f32);
    } else {
      b.extend_from_slice(&[0u8; 20]);
    }
  }
  b
}

// RnaContraTables = FoldScoreSets::new(0.).transfer() (src/mccaskill_algo.rs:25-210), field for field
fn contra_blob(s: &FoldScoreSets) -> Vec<u8> {
  let mut b = Vec::new();
  put_i32(&mut b, MAX_LOOP_LEN as i32);
  put_i32(&mut b, MIN_SPAN_HAIRPIN_CLOSE as i32);
  put_i32(&mut b, MAX_INTERIOR_EXPLICIT as i32);
  put_i32(&mut b, 0);
  let flat1 = |b: &mut Vec<u8>, t: &[f32]| { for x in t.iter() { put_f32(b, *x); } };
  flat1(&mut b, &s.hairpin_scores_len[..]);
  flat1(&mut b, &s.bulge_scores_len[..]);
  flat1(&mut b, &s.interior_scores_len[..]);
  flat1(&mut b, &s.interior_scores_symmetric[..]);
  flat1(&mut b, &s.interior_scores_asymmetric[..]);
  for t in [&s.stack_scores, &s.terminal_mismatch_scores].iter() {
    for w in t.iter() { for x in w.iter() { for y in x.iter() { for z in y.iter() { put_f32(&mut b, *z); } } } }
  }
  for t in [&s.dangling_scores_left, &s.dangling_scores_right].iter() {
    for w in t.iter() { for x in w.iter() { for y in x.iter() { put_f32(&mut b, *y); } } }
  }
  for t in [&s.helix_close_scores, &s.basepair_scores].iter() { for w in t.iter() { for x in w.iter() { put_f32(&mut b, *x); } } }
  for w in s.interior_scores_explicit.iter() { for x in w.iter() { put_f32(&mut b, *x); } }
  flat1(&mut b, &s.bulge_scores_0x1[..]);
  for w in s.interior_scores_1x1.iter() { for x in w.iter() { put_f32(&mut b, *x); } }
  for x in [s.multibranch_score_base, s.multibranch_score_basepair, s.multibranch_score_unpair,
            s.external_score_basepair, s.external_score_unpair].iter() { put_f32(&mut b, *x); }
  flat1(&mut b, &s.hairpin_scores_len_cumulative[..]);
  flat1(&mut b, &s.bulge_scores_len_cumulative[..]);
  flat1(&mut b, &s.interior_scores_len_cumulative[..]);
  flat1(&mut b, &s.interior_scores_symmetric_cumulative[..]);
  flat1(&mut b, &s.interior_scores_asymmetric_cumulative[..]);
  b
}

fn packed_bpp(bpp: &SparseProbMat<u8>, len: usize) -> Vec<f32> {
  let mut v = vec![-1.0f32; len * (len - 1) / 2];   // RNA_BPP_ABSENT
  for (&(i, j), &p) in bpp.iter() {
    let (i, j) = (i as usize, j as usize);
    v[i * (2 * len - i - 1) / 2 + (j - i - 1)] = p;
  }
  v
}

fn main() {
  let args: Vec<String> = std::env::args().collect();
  if args.len() != 3 { eprintln!("usage: ref_dump <assets/sampled_trnas.fa> <out dir>"); std::process::exit(2); }
  let out = Path::new(&args[2]);
  create_dir_all(out).unwrap();
  let mut fold_score_sets = FoldScoreSets::new(0.);
  fold_score_sets.transfer();
  write_blob(&out.join("turner2004.tbl"), 1, &turner_blob());
  write_blob(&out.join("contrafold_v202.tbl"), 2, &contra_blob(&fold_score_sets));

  let reader = bio::io::fasta::Reader::from_file(Path::new(&args[1])).unwrap();
  let seqs: Vec<Seq> = reader.records().map(|r| bytes2seq(r.unwrap().seq())).collect();
  let gammas: Vec<f32> = (-7..11).map(|p| (2. as f32).powi(p)).collect();   // src/bin/centroid_fold.rs:9-11,148-149
  for (s, seq) in seqs.iter().enumerate() {
    let len = seq.len();
    for &contra in [false, true].iter() {
      let tag = format!("seq{}_{}", s, if contra { "contra" } else { "turner" });
      let (bpp, _) = mccaskill_algo::<u8>(&seq[..], contra, false, &fold_score_sets);
      let mut fold_scores = FoldScores::<u8>::new();
      let sums = if contra { get_fold_sums_contra::<u8>(&seq[..], &mut fold_scores, false, &fold_score_sets) }
                 else { get_fold_sums::<u8>(&seq[..], &mut fold_scores) };
      write_npy(&out.join(format!("{}_logz.npy", tag)), "<f4", &[1], &f32s(&[sums.sums_external[0][len - 1]]));
      write_npy(&out.join(format!("{}_bpp.npy", tag)), "<f4", &[len * (len - 1) / 2], &f32s(&packed_bpp(&bpp, len)));
      let mut structs: Vec<u8> = Vec::new();
      let mut eas: Vec<f32> = Vec::new();
      for &g in gammas.iter() {
        let cf = centroid_fold::<u8>(&bpp, len, g);
        let mut st = vec![b'.'; len];
        for &(i, j) in cf.basepair_pos_pairs.iter() { st[i as usize] = b'('; st[j as usize] = b')'; }
        structs.extend_from_slice(&st);
        eas.push(cf.expect_accuracy);
      }
      write_npy(&out.join(format!("{}_structs.npy", tag)), "|u1", &[gammas.len(), len], &structs);
      write_npy(&out.join(format!("{}_expect_acc.npy", tag)), "<f4", &[gammas.len()], &f32s(&eas));
    }
  }
  // durbin_algo on all pairs of sentinel-padded sequences (src/bin/durbin_algo.rs:44-75, tests/tests.rs:45-80)
  let mut align_scores = AlignScores::new(0.);
  align_scores.transfer();
  let padded: Vec<Seq> = seqs.iter().map(|s| { let mut p = s.clone(); p.insert(0, PSEUDO_BASE); p.push(PSEUDO_BASE); p }).collect();
  for a in 0..padded.len() {
    for b in a + 1..padded.len() {
      let m = durbin_algo(&(&padded[a][..], &padded[b][..]), &align_scores);
      let flat: Vec<f32> = m.iter().flat_map(|r| r.iter().cloned()).collect();
      write_npy(&out.join(format!("durbin_{}_{}.npy", a, b)), "<f4", &[padded[a].len(), padded[b].len()], &f32s(&flat));
    }
  }
  write_npy(&out.join("gammas.npy"), "<f4", &[gammas.len()], &f32s(&gammas));
}
