"""Real-data ingestion for BASELINE configs[2] / [4]: Stockholm alignments -> families of
ungapped sequences, with the reference's own reading rules and its scripts' family filter, plus the accuracy counts
of its evaluation script.  Host-side only (no GPU work here).

reference                                                         here
----------------------------------------------------------------  ------------------------------------------
utils::align_char2base                src/utils.rs:746-754         align_char2base
utils::read_align_stockholm           src/utils.rs:719-744         read_align_stockholm
scripts/compile_rna_fams.py:32-34     family filter                iter_stockholm_families + compile_rna_fams
scripts/compile_rna_fams.py:66-70     is_valid (no IUPAC codes)    is_valid
scripts/compile_rna_fams.py:72-109    convert_css / recover_ss     convert_css, recover_ss
scripts/get_stats_of_ss_estimation_programs.py:154-197             pos_neg_counts, ppv / sens / f1 / mcc

`assets/rfam_seed_stas_v14.3.sth` is not in the reference tree (.MISSING_LARGE_BLOBS); when it is supplied,
`compile_rna_fams(path)` yields the families bench.py's --rfam-sth option folds and aligns instead of the seeded
synthetic stand-in families."""
from __future__ import annotations

from math import sqrt
from typing import Dict, Iterator, List, Sequence, Tuple

import numpy as np

PSEUDO_BASE = 4
_BASE = {"a": 0, "A": 0, "c": 1, "C": 1, "g": 2, "G": 2, "u": 3, "U": 3}
BRACKET_PAIRS = [("(", ")"), ("A", "a"), ("B", "b"), ("C", "c"), ("D", "d"), ("E", "e")]


def align_char2base(x: str) -> int:
    """src/utils.rs:746-754: a/c/g/u in either case, everything else (gaps included) is PSEUDO_BASE."""
    return _BASE.get(x, PSEUDO_BASE)


def read_align_stockholm(path: str) -> Tuple[List[List[int]], List[str]]:
    """src/utils.rs:719-744, rule for rule: empty lines and lines starting with '#' are skipped, '//' ends the
    alignment, every other line is `id sequence` (one row per line: no interleaved blocks); returns the alignment
    COLUMN-major (`Cols`) and the row ids."""
    seq_ids: List[str] = []
    seqs: List[List[int]] = []
    with open(path) as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line == "" or line.startswith("#"):
                continue
            if line.startswith("//"):
                break
            parts = line.split()
            seq_ids.append(parts[0])
            seqs.append([align_char2base(c) for c in parts[1]])
    if not seqs:
        raise ValueError(f"{path}: no alignment rows")   # (the reference panics on seqs[0])
    align_len = len(seqs[0])
    cols = [[s[i] for s in seqs] for i in range(align_len)]
    return cols, seq_ids


def iter_stockholm_families(path: str) -> Iterator[Dict]:
    """Every alignment of a multi-alignment Stockholm file (Rfam.seed style), interleaved blocks joined per id, '.'
    gaps read as '-' (what Bio.AlignIO's Stockholm parser hands scripts/compile_rna_fams.py): dicts with `ids`, `rows`
    (gapped strings), `ss_cons` (the #=GC SS_cons line) and `accession`."""
    ids: List[str] = []
    rows: Dict[str, List[str]] = {}
    ss: List[str] = []
    acc = ""
    # Rfam seed files are Latin-1 (author names)
    with open(path, encoding="latin-1") as fh:
        for line in fh:
            line = line.rstrip("\n")
            if line.startswith("//"):
                if ids:
                    yield dict(ids=ids, rows=["".join(rows[i]) for i in ids], ss_cons="".join(ss), accession=acc)
                ids, rows, ss, acc = [], {}, [], ""
                continue
            if line == "" or line.startswith("# STOCKHOLM"):
                continue
            if line.startswith("#=GC SS_cons"):
                ss.append(line.split()[2])
                continue
            if line.startswith("#=GF AC"):
                acc = line.split()[2]
                continue
            if line.startswith("#"):
                continue
            parts = line.split()
            if len(parts) < 2:
                continue
            if parts[0] not in rows:
                ids.append(parts[0])
                rows[parts[0]] = []
            rows[parts[0]].append(parts[1].replace(".", "-"))
    if ids:
        yield dict(ids=ids, rows=["".join(rows[i]) for i in ids], ss_cons="".join(ss), accession=acc)


def is_valid(rows: Sequence[str]) -> bool:
    """scripts/compile_rna_fams.py:66-70: no IUPAC ambiguity code (upper case, as there) in any row."""
    return not any(ch in row for row in rows for ch in "RYWSMKHBVDN")


def convert_css(css: str) -> str:
    """scripts/compile_rna_fams.py:72-83"""
    out = []
    for ch in css:
        if ch in "(<[{":
            out.append("(")
        elif ch in ")>]}":
            out.append(")")
        elif ch in "ABCDEabcde":
            out.append(ch)
        else:
            out.append(".")
    return "".join(out)


def recover_ss(css: str, seq_with_gaps: str) -> str:
    """scripts/compile_rna_fams.py:85-109: the consensus structure projected onto one ungapped row."""
    pos_map, pos = {}, 0
    for i, ch in enumerate(seq_with_gaps):
        if ch != "-":
            pos_map[i] = pos
            pos += 1
    rec = ["."] * pos
    for left, right in BRACKET_PAIRS:
        stack: List[int] = []
        for i, ch in enumerate(css):
            if ch == left:
                stack.append(i)
            elif ch == right:
                j = stack.pop()
                if seq_with_gaps[j] == "-" or seq_with_gaps[i] == "-":
                    continue
                rec[pos_map[j]] = left
                rec[pos_map[i]] = right
    return "".join(rec)


def compile_rna_fams(path: str, max_seq_num: int = 10, max_sa_len: int = 200) -> List[Dict]:
    """The family filter of scripts/compile_rna_fams.py:32-34 — at most `max_seq_num` rows, at most `max_sa_len`
    alignment columns, no IUPAC codes — and per row the ungapped sequence (base codes 0..3; rows with any other
    character, which `bytes2seq` would panic on, drop the family) and its reference structure."""
    fams = []
    for f in iter_stockholm_families(path):
        rows = f["rows"]
        if len(rows) > max_seq_num or len(rows[0]) > max_sa_len or not is_valid(rows):
            continue
        css = convert_css(f["ss_cons"]) if f["ss_cons"] else "." * len(rows[0])
        seqs, sss, ok = [], [], True
        for row in rows:
            ungapped = row.replace("-", "")
            codes = [_BASE.get(c, -1) for c in ungapped]
            if not ungapped or min(codes) < 0:
                ok = False
                break
            seqs.append(np.array(codes, dtype=np.uint8))
            sss.append(recover_ss(css, row))
        if ok:
            fams.append(dict(accession=f["accession"], ids=f["ids"], seqs=seqs, ref_sss=sss))
    return fams


def intra_family_pairs(fams: Sequence[Dict]) -> Tuple[List[np.ndarray], np.ndarray]:
    """All sequences of the families, concatenated, and every pair (a < b) WITHIN a family — BASELINE configs[4]."""
    seqs: List[np.ndarray] = []
    pairs = []
    for f in fams:
        base = len(seqs)
        n = len(f["seqs"])
        seqs.extend(f["seqs"])
        pairs.extend((base + a, base + b) for a in range(n) for b in range(a + 1, n))
    return seqs, np.array(pairs, dtype=np.uint32).reshape(-1, 2)


def ss_pairs(ss: str) -> set:
    """Base pairs of a (pseudo-knotted) bracket string, bracket kinds as in BRACKET_PAIRS."""
    out = set()
    for left, right in BRACKET_PAIRS:
        stack: List[int] = []
        for i, ch in enumerate(ss):
            if ch == left:
                stack.append(i)
            elif ch == right and stack:
                out.add((stack.pop(), i))
    return out


def pos_neg_counts(estimated_sss: Sequence[str], ref_sss: Sequence[str]) -> Tuple[int, int, int, int]:
    """scripts/get_stats_of_ss_estimation_programs.py:154-173: (tp, tn, fp, fn) over all position pairs i < j."""
    tp = tn = fp = fn = 0
    for est, ref in zip(estimated_sss, ref_sss):
        L = len(ref)
        e, r = ss_pairs(est), ss_pairs(ref)
        both = len(e & r)
        tp += both
        fp += len(e) - both
        fn += len(r) - both
        tn += L * (L - 1) // 2 - len(e | r)
    return tp, tn, fp, fn


def get_ppv(tp, fp): return tp / (tp + fp)
def get_sens(tp, fn): return tp / (tp + fn)
def get_fpr(tn, fp): return fp / (tn + fp)
def get_f1_score(ppv, sens): return 2 * ppv * sens / (ppv + sens)
def get_mcc(tp, tn, fp, fn): return (tp * tn - fp * fn) / sqrt((tp + fp) * (tp + fn) * (tn + fp) * (tn + fn))
