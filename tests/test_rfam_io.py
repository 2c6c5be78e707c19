"""N4: the Stockholm reader's rules (src/utils.rs:719-754), the family filter and structure projection of
scripts/compile_rna_fams.py, and the accuracy counts of scripts/get_stats_of_ss_estimation_programs.py.  CPU only."""
import numpy as np

from rna_algos_b200 import rfam

STH = """# STOCKHOLM 1.0
#=GF AC   RF00001
#=GF AU   Someone
seq1/1-10    GGGA.AUCCC
seq2/5-14    GGCAAAGCC-
#=GC SS_cons <<<....>>>

seq1/1-10    AC
seq2/5-14    -C
#=GC SS_cons ..
//
# STOCKHOLM 1.0
#=GF AC   RF00002
a            ACGUN
b            ACGUA
#=GC SS_cons .....
//
# STOCKHOLM 1.0
#=GF AC   RF00003
x            ACGUACGUAC
#=GC SS_cons (((....)))
//
"""


def test_align_char2base_and_single_alignment_reader(tmp_path):
    assert [rfam.align_char2base(c) for c in "aAcCgGuU-.TN"] == [0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 4, 4]
    p = tmp_path / "one.sth"
    p.write_text("# STOCKHOLM 1.0\n\n#=GF ID x\nr1  AC-GU\nr2  acNgu\n//\nr3  AAAAA\n")
    cols, ids = rfam.read_align_stockholm(str(p))
    assert ids == ["r1", "r2"]                      # '//' ends the alignment
    assert cols == [[0, 0], [1, 1], [4, 4], [2, 2], [3, 3]]   # column-major, gap and N are PSEUDO_BASE


def test_family_filter_and_projection(tmp_path):
    p = tmp_path / "seed.sth"
    p.write_text(STH)
    fams = list(rfam.iter_stockholm_families(str(p)))
    assert [f["accession"] for f in fams] == ["RF00001", "RF00002", "RF00003"]
    assert fams[0]["rows"] == ["GGGA-AUCCCAC", "GGCAAAGCC--C"] and fams[0]["ss_cons"] == "<<<....>>>.."
    kept = rfam.compile_rna_fams(str(p))
    assert [f["accession"] for f in kept] == ["RF00001", "RF00003"]     # RF00002 has an IUPAC code
    f = kept[0]
    assert ["".join("ACGU"[b] for b in s) for s in f["seqs"]] == ["GGGAAUCCCAC", "GGCAAAGCCC"]
    assert f["ref_sss"][0] == "(((....)))..".replace("....", "...")       # the gap column of row 1 is gone
    # row 2: the pair whose right partner is a gap is dropped
    assert f["ref_sss"][1] == ".((....))."
    assert [x["accession"] for x in rfam.compile_rna_fams(str(p), max_sa_len=10)] == ["RF00003"]
    seqs, pairs = rfam.intra_family_pairs(kept)
    assert len(seqs) == 3 and pairs.tolist() == [[0, 1]]


def test_accuracy_counts():
    ref = ["(((...)))", "..((..)).."]
    est = ["((.....))", "..((..)).."]
    tp, tn, fp, fn = rfam.pos_neg_counts(est, ref)
    assert (tp, fp, fn) == (4, 0, 1) and tn == 9 * 8 // 2 + 10 * 9 // 2 - 5
    assert abs(rfam.get_ppv(tp, fp) - 1.0) < 1e-12 and abs(rfam.get_sens(tp, fn) - 0.8) < 1e-12
    assert abs(rfam.get_f1_score(1.0, 0.8) - 8 / 9) < 1e-12
    assert 0.8 < rfam.get_mcc(tp, tn, fp, fn) < 1.0
    assert rfam.ss_pairs("(A.a)") == {(0, 4), (1, 3)}
