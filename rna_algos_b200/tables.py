"""Table blobs of include/rna_algos_b200.h as ctypes structures, plus table constructors.

The reference takes its Turner 2004 / CONTRAfold v2.02 numbers from the crates.io dependency
``rna-ss-params = "0.1"`` (reference Cargo.toml:12, glob-imported at src/utils.rs:8-10), which is NOT
vendored in the reference tree and is not available offline.  The blobs are therefore RUNTIME
arguments of the C ABI; a Rust shim fills them from the genuine crate.  What this
module offers:

* ``contralign_tables()``   – the genuine CONTRAlign v2.01 numbers (in-tree: src/compiled_align_scores.rs:2-19).
* ``standin_turner_tables()`` – a STAND-IN of the shape of the Turner 2004 set, NOT the upstream numbers: stacking,
  loop-initiation, multiloop, Ninio, terminal-AU and a few special hairpins are restated from the public
  nearest-neighbour parameters from memory; mismatch / dangle / 1x1 / 1x2 / 2x2 tables are deterministic
  pseudo-values.  Results computed with it differ from the reference's.  It exists so that kernels,
  oracle and benchmarks have a realistic table set to run on.
* ``standin_contra_tables()`` – same status for CONTRAfold v2.02, materialised exactly like
  FoldScoreSets::new(0.).transfer() (src/mccaskill_algo.rs:25-210): only canonical entries are written,
  everything else keeps init_val = 0.
* ``load_table_file()`` / ``load_genuine_tables(dir)`` – read ``turner2004.tbl`` / ``contrafold_v202.tbl`` blobs
  written from the genuine crate by ``tools/ref_dump`` (a machine with cargo); these names are reserved for
  genuine values — the stand-ins are dumped as ``standin_turner.tbl`` / ``standin_contrafold.tbl``.
* ``random_*_tables(seed)`` – fuzz tables for bit-parity tests of kernels vs oracle.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

NUM_BASES = 4
A, Cb, G, U = 0, 1, 2, 3
LOOP_TABLE_LEN = 31
MAX_SPECIAL_HAIRPINS = 128
MAX_SPECIAL_HAIRPIN_LEN = 12
CONTRA_MAX_LOOP_LEN = 30
CONTRA_MAX_INTERIOR_SYMMETRIC = 15
CONTRA_MAX_INTERIOR_ASYMMETRIC = 28
CONTRA_MAX_INTERIOR_EXPLICIT = 4
BPP_ABSENT = -1.0

CANONICAL = {(A, U), (Cb, G), (G, Cb), (G, U), (U, A), (U, G)}
AUGU = {(A, U), (U, A), (G, U), (U, G)}

f32 = C.c_float


def _arr(*dims):
    t = f32
    for d in reversed(dims):
        t = t * d
    return t


class SpecialHairpin(C.Structure):
    _fields_ = [
        ("len", C.c_uint8),
        ("seq", C.c_uint8 * MAX_SPECIAL_HAIRPIN_LEN),
        ("_pad", C.c_uint8 * 3),
        ("score", f32),
    ]


class TurnerTables(C.Structure):
    """RnaTurnerTables (include/rna_algos_b200.h)."""

    _fields_ = [
        ("max_2loop_len", C.c_int32),
        ("min_span_hairpin_close", C.c_int32),
        ("min_hairpin_len", C.c_int32),
        ("max_hairpin_len_extrapolation", C.c_int32),
        ("min_hairpin_len_extrapolation", C.c_int32),
        ("num_special_hairpins", C.c_int32),
        ("coeff_hairpin_len_extrapolation", f32),
        ("helix_augu_end_penalty", f32),
        ("ninio_coeff", f32),
        ("ninio_max", f32),
        ("init_multibranch_base", f32),
        ("coeff_num_branches", f32),
        ("hairpin_scores_init", _arr(LOOP_TABLE_LEN)),
        ("bulge_scores_init", _arr(LOOP_TABLE_LEN)),
        ("interior_scores_init", _arr(LOOP_TABLE_LEN)),
        ("stack_scores", _arr(4, 4, 4, 4)),
        ("terminal_mismatch_scores_hairpin", _arr(4, 4, 4, 4)),
        ("terminal_mismatch_scores_1xmany", _arr(4, 4, 4, 4)),
        ("terminal_mismatch_scores_2x3", _arr(4, 4, 4, 4)),
        ("terminal_mismatch_scores_interior", _arr(4, 4, 4, 4)),
        ("terminal_mismatch_scores_multibranch", _arr(4, 4, 4, 4)),
        ("dangling_scores_5prime", _arr(4, 4, 4)),
        ("dangling_scores_3prime", _arr(4, 4, 4)),
        ("interior_scores_1x1", _arr(4, 4, 4, 4, 4, 4)),
        ("interior_scores_1x2", _arr(4, 4, 4, 4, 4, 4, 4)),
        ("interior_scores_2x2", _arr(4, 4, 4, 4, 4, 4, 4, 4)),
        ("hairpin_scores_special", SpecialHairpin * MAX_SPECIAL_HAIRPINS),
    ]


class ContraTables(C.Structure):
    """RnaContraTables == FoldScoreSets (reference src/utils.rs:91-119) + caps."""

    _fields_ = [
        ("max_loop_len", C.c_int32),
        ("min_span_hairpin_close", C.c_int32),
        ("max_interior_explicit", C.c_int32),
        ("_pad", C.c_int32),
        ("hairpin_scores_len", _arr(CONTRA_MAX_LOOP_LEN + 1)),
        ("bulge_scores_len", _arr(CONTRA_MAX_LOOP_LEN)),
        ("interior_scores_len", _arr(CONTRA_MAX_LOOP_LEN - 1)),
        ("interior_scores_symmetric", _arr(CONTRA_MAX_INTERIOR_SYMMETRIC)),
        ("interior_scores_asymmetric", _arr(CONTRA_MAX_INTERIOR_ASYMMETRIC)),
        ("stack_scores", _arr(4, 4, 4, 4)),
        ("terminal_mismatch_scores", _arr(4, 4, 4, 4)),
        ("dangling_scores_left", _arr(4, 4, 4)),
        ("dangling_scores_right", _arr(4, 4, 4)),
        ("helix_close_scores", _arr(4, 4)),
        ("basepair_scores", _arr(4, 4)),
        ("interior_scores_explicit", _arr(CONTRA_MAX_INTERIOR_EXPLICIT, CONTRA_MAX_INTERIOR_EXPLICIT)),
        ("bulge_scores_0x1", _arr(4)),
        ("interior_scores_1x1", _arr(4, 4)),
        ("multibranch_score_base", f32),
        ("multibranch_score_basepair", f32),
        ("multibranch_score_unpair", f32),
        ("external_score_basepair", f32),
        ("external_score_unpair", f32),
        ("hairpin_scores_len_cumulative", _arr(CONTRA_MAX_LOOP_LEN + 1)),
        ("bulge_scores_len_cumulative", _arr(CONTRA_MAX_LOOP_LEN)),
        ("interior_scores_len_cumulative", _arr(CONTRA_MAX_LOOP_LEN - 1)),
        ("interior_scores_symmetric_cumulative", _arr(CONTRA_MAX_INTERIOR_SYMMETRIC)),
        ("interior_scores_asymmetric_cumulative", _arr(CONTRA_MAX_INTERIOR_ASYMMETRIC)),
    ]


class AlignTables(C.Structure):
    """RnaAlignTables == AlignScores (reference src/durbin_algo.rs:4-14)."""

    _fields_ = [
        ("match2match_score", f32),
        ("match2insert_score", f32),
        ("insert_extend_score", f32),
        ("insert_switch_score", f32),
        ("init_match_score", f32),
        ("init_insert_score", f32),
        ("insert_scores", _arr(4)),
        ("match_scores", _arr(4, 4)),
    ]


def view(struct, field) -> np.ndarray:
    """Writable float32 numpy view of an array field of a ctypes table struct."""
    return np.ctypeslib.as_array(getattr(struct, field))


# ----------------------------------------------------------------------------------------------
# CONTRAlign v2.01 — genuine values (reference src/compiled_align_scores.rs:2-19)
# ----------------------------------------------------------------------------------------------
def contralign_tables() -> AlignTables:
    t = AlignTables()
    m = [
        [0.5256508867, -0.40906402, -0.2502759109, -0.3252306723],
        [-0.40906402, 0.6665219366, -0.3289391181, -0.1326088918],
        [-0.2502759109, -0.3289391181, 0.6684676551, -0.3565888168],
        [-0.3252306723, -0.1326088918, -0.3565888168, 0.459052045],
    ]
    view(t, "match_scores")[...] = np.array(m, dtype=np.float32)
    view(t, "insert_scores")[...] = np.array(
        [-0.002521927159, -0.08313891561, -0.07443970653, -0.01290054598], dtype=np.float32
    )
    t.init_match_score = 0.3959924457
    t.init_insert_score = -0.3488104904
    t.match2match_score = 2.50575671
    t.match2insert_score = 0.1970448791
    t.insert_extend_score = 1.014026583
    t.insert_switch_score = -7.346968782
    return t


# ----------------------------------------------------------------------------------------------
# Turner 2004 — restated / stand-in (see module docstring)
# ----------------------------------------------------------------------------------------------
_KT = 1.98717e-3 * 310.15  # kcal/mol at 37 C; scores are -dG/kT (log-weights)


def _sc(kcal):
    return np.float32(-np.float64(kcal) / _KT)


# pair-type order of the public parameter files: CG GC GU UG AU UA
_PT = [(Cb, G), (G, Cb), (G, U), (U, G), (A, U), (U, A)]
_STACK_KCAL = [  # stack[type(i,j)][type(l,k)], kcal/mol
    [-2.4, -3.3, -2.1, -1.4, -2.1, -2.1],
    [-3.3, -3.4, -2.5, -1.5, -2.2, -2.4],
    [-2.1, -2.5, 1.3, -0.5, -1.4, -1.3],
    [-1.4, -1.5, -0.5, 0.3, -0.6, -1.0],
    [-2.1, -2.2, -1.4, -0.6, -1.1, -0.9],
    [-2.1, -2.4, -1.3, -1.0, -0.9, -1.3],
]
_HAIRPIN_KCAL = [np.inf, np.inf, np.inf, 5.4, 5.6, 5.7, 5.4, 6.0, 5.5, 6.4, 6.5, 6.6, 6.7, 6.8, 6.9, 6.9,
                 7.0, 7.1, 7.1, 7.2, 7.2, 7.3, 7.3, 7.4, 7.4, 7.5, 7.5, 7.5, 7.6, 7.6, 7.7]
_BULGE_KCAL = [np.inf, 3.8, 2.8, 3.2, 3.6, 4.0, 4.4, 4.59, 4.7, 4.8, 4.9, 5.0, 5.1, 5.2, 5.3, 5.4, 5.4,
               5.5, 5.5, 5.6, 5.7, 5.7, 5.8, 5.8, 5.8, 5.9, 5.9, 6.0, 6.0, 6.0, 6.1]
_INTERIOR_KCAL = [np.inf, np.inf, 1.0, 1.0, 1.1, 2.0, 2.0, 2.1, 2.3, 2.4, 2.5, 2.6, 2.7, 2.8, 2.9, 2.9,
                  3.0, 3.1, 3.1, 3.2, 3.3, 3.3, 3.4, 3.4, 3.5, 3.5, 3.5, 3.6, 3.6, 3.7, 3.7]
_SPECIAL = [  # (loop incl. closing pair, kcal/mol)
    ("CAACG", 6.8), ("GUUAC", 6.9),
    ("CAACGG", 5.5), ("CCAAGG", 3.3), ("CCACGG", 3.7), ("CCCAGG", 3.4), ("CCGAGG", 3.5),
    ("CCGCGG", 3.6), ("CCUAGG", 3.7), ("CCUCGG", 2.5), ("CUAAGG", 3.6), ("CUACGG", 2.8),
    ("CUCAGG", 3.7), ("CUCCGG", 2.7), ("CUGCGG", 2.8), ("CUUAGG", 3.5), ("CUUCGG", 3.7),
    ("CUUUGG", 3.7), ("CGAAAG", 3.0), ("GGGGAC", 3.0), ("GGUGAC", 3.0), ("CGAGAG", 3.0),
    ("GGAGAC", 3.0), ("CGCAAG", 3.0), ("GGAAAC", 3.0), ("CGGAAG", 3.0), ("CUUCGG", 3.0),
    ("CGUGAG", 3.0), ("CGAAGG", 2.5), ("CUACGG", 3.0), ("GGCAAC", 3.0), ("CGCGAG", 3.0),
    ("UGAGAG", 3.0), ("CGAGAG", 3.0), ("AGAAAU", 3.0), ("CGUAAG", 3.0), ("CUAACG", 3.0),
    ("UGAAAG", 3.0), ("GGAAGC", 3.0), ("GGGAAC", 3.0), ("UGAAAA", 3.0), ("AGCAAU", 3.0),
    ("AGUAAU", 3.0), ("CGGGAG", 3.0), ("AGUGAU", 3.0), ("GGCGAC", 3.0), ("GGGAGC", 3.0),
    ("GUGAAC", 3.0), ("UGGAAA", 3.0),
    ("ACAGUACU", 2.8), ("ACAGUGAU", 3.6), ("ACAGUGCU", 2.9), ("ACAGUGUU", 1.8),
]
_B = {"A": A, "C": Cb, "G": G, "U": U}


def _hash01(*xs) -> float:
    """Deterministic pseudo-random in [0,1) from small integers (splitmix64-style)."""
    z = 0x9E3779B97F4A7C15
    for x in xs:
        z = (z ^ (int(x) + 0x9E3779B97F4A7C15 + ((z << 6) & 0xFFFFFFFFFFFFFFFF) + (z >> 2))) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
        z ^= z >> 31
    return (z >> 11) / float(1 << 53)


def _pair_strength(x, y) -> float:
    """Rough closing-pair stability in kcal/mol (stand-in tables only)."""
    if (x, y) in ((Cb, G), (G, Cb)):
        return 1.0
    if (x, y) in ((A, U), (U, A)):
        return 0.55
    if (x, y) in ((G, U), (U, G)):
        return 0.45
    return 0.2


def standin_turner_tables() -> TurnerTables:
    t = TurnerTables()
    t.max_2loop_len = 30
    t.min_span_hairpin_close = 5
    t.min_hairpin_len = 3
    t.max_hairpin_len_extrapolation = 9
    t.min_hairpin_len_extrapolation = 10
    t.coeff_hairpin_len_extrapolation = float(_sc(1.75 * _KT))  # -(1.75 RT)/RT = -1.75
    t.helix_augu_end_penalty = float(_sc(0.5))
    t.ninio_coeff = float(_sc(0.6))
    t.ninio_max = float(_sc(3.0))
    t.init_multibranch_base = float(_sc(9.3))
    t.coeff_num_branches = float(_sc(-0.9))
    big = np.float32(-1.0e4)  # stands for "forbidden" lengths; never indexed by the recurrences
    for name, kc in (("hairpin_scores_init", _HAIRPIN_KCAL), ("bulge_scores_init", _BULGE_KCAL),
                     ("interior_scores_init", _INTERIOR_KCAL)):
        v = view(t, name)
        for n, e in enumerate(kc):
            v[n] = big if np.isinf(e) else _sc(e)
    st = view(t, "stack_scores")
    st[...] = 0.0
    for a, (i, j) in enumerate(_PT):
        for b, (l, k) in enumerate(_PT):
            st[i, j, k, l] = _sc(_STACK_KCAL[a][b])
    # stand-in mismatch / dangle tables: stabilising, scaled by closing-pair strength
    for fi, (name, lo, hi) in enumerate((
        ("terminal_mismatch_scores_hairpin", -1.6, -0.3),
        ("terminal_mismatch_scores_1xmany", -0.4, 0.7),
        ("terminal_mismatch_scores_2x3", -1.2, 0.7),
        ("terminal_mismatch_scores_interior", -1.1, 0.7),
        ("terminal_mismatch_scores_multibranch", -1.5, -0.1),
    )):
        v = view(t, name)
        for i in range(4):
            for j in range(4):
                for x in range(4):
                    for y in range(4):
                        e = lo + (hi - lo) * _hash01(11 + fi, i, j, x, y)
                        e *= 0.5 + 0.5 * _pair_strength(i, j)
                        v[i, j, x, y] = _sc(e)
    for fi, (name, lo, hi) in enumerate((("dangling_scores_5prime", -0.5, -0.1),
                                         ("dangling_scores_3prime", -1.7, -0.1))):
        v = view(t, name)
        for i in range(4):
            for j in range(4):
                for x in range(4):
                    e = lo + (hi - lo) * _hash01(31 + fi, i, j, x)
                    v[i, j, x] = _sc(e * (0.5 + 0.5 * _pair_strength(i, j)))
    i11 = view(t, "interior_scores_1x1")
    i12 = view(t, "interior_scores_1x2")
    i22 = view(t, "interior_scores_2x2")
    idx = np.indices((4,) * 6)
    h = np.vectorize(lambda *xs: _hash01(41, *xs))(*idx)
    i11[...] = (-(0.4 + 1.5 * h) / _KT).astype(np.float32)
    idx = np.indices((4,) * 7)
    h = np.vectorize(lambda *xs: _hash01(42, *xs))(*idx)
    i12[...] = (-(2.2 + 1.8 * h) / _KT).astype(np.float32)
    rng = np.random.default_rng(20041)  # 4^8 entries: vectorised generator instead of the scalar hash
    i22[...] = (-(0.5 + 2.5 * rng.random((4,) * 8)) / _KT).astype(np.float32)
    seen = []
    for s, e in _SPECIAL:
        if s in seen:
            continue
        seen.append(s)
        ent = t.hairpin_scores_special[len(seen) - 1]
        ent.len = len(s)
        for p, ch in enumerate(s):
            ent.seq[p] = _B[ch]
        ent.score = float(_sc(e))
    t.num_special_hairpins = len(seen)
    return t


# ----------------------------------------------------------------------------------------------
# CONTRAfold v2.02 — restated scalars / stand-in tables, materialised like FoldScoreSets::transfer
# ----------------------------------------------------------------------------------------------
def standin_contra_tables(lib=None) -> ContraTables:
    t = ContraTables()  # FoldScoreSets::new(0.)
    t.max_loop_len = CONTRA_MAX_LOOP_LEN
    t.min_span_hairpin_close = 5
    t.max_interior_explicit = CONTRA_MAX_INTERIOR_EXPLICIT
    hp = [-5.993180158, -3.108105762, 0.4468046273, 2.105729061, 1.823291902, -0.09468101504]
    v = view(t, "hairpin_scores_len")
    for n in range(CONTRA_MAX_LOOP_LEN + 1):
        v[n] = hp[n] if n < len(hp) else -0.35 + 0.5 * _hash01(51, n)
    v = view(t, "bulge_scores_len")
    for n in range(CONTRA_MAX_LOOP_LEN):
        v[n] = -2.4 if n == 0 else -0.6 + 0.7 * _hash01(52, n)
    v = view(t, "interior_scores_len")
    for n in range(CONTRA_MAX_LOOP_LEN - 1):
        v[n] = -0.43 if n == 0 else -0.45 + 0.5 * _hash01(53, n)
    v = view(t, "interior_scores_symmetric")
    for n in range(CONTRA_MAX_INTERIOR_SYMMETRIC):
        v[n] = -0.4 + 0.7 * _hash01(54, n)
    v = view(t, "interior_scores_asymmetric")
    for n in range(CONTRA_MAX_INTERIOR_ASYMMETRIC):
        v[n] = -2.1 if n == 0 else -0.5 + 0.5 * _hash01(55, n)
    bp = {(A, U): 0.59791199, (U, A): 0.59791199, (Cb, G): 1.544290641, (G, Cb): 1.544290641,
          (G, U): -0.01304754992, (U, G): -0.01304754992}
    st = view(t, "stack_scores")
    tm = view(t, "terminal_mismatch_scores")
    dl = view(t, "dangling_scores_left")
    dr = view(t, "dangling_scores_right")
    hc = view(t, "helix_close_scores")
    bs = view(t, "basepair_scores")
    for (i, j) in sorted(CANONICAL):
        for (k, l) in sorted(CANONICAL):
            # helix_stacking is shared between the two read directions in CONTRAfold
            key = min((i, j, k, l), (l, k, j, i))
            st[i, j, k, l] = 0.1 + 1.2 * _hash01(61, *key) * (_pair_strength(i, j) + _pair_strength(k, l)) / 2.0
        for x in range(4):
            for y in range(4):
                tm[i, j, x, y] = -0.9 + 1.4 * _hash01(62, i, j, x, y)
            dl[i, j, x] = -0.3 + 0.5 * _hash01(63, i, j, x)
            dr[i, j, x] = -0.3 + 0.6 * _hash01(64, i, j, x)
        hc[i, j] = -0.98 + 0.6 * _pair_strength(i, j) + 0.1 * _hash01(65, min(i, j), max(i, j))
        bs[i, j] = bp[(i, j)]
    ex = view(t, "interior_scores_explicit")
    for a in range(CONTRA_MAX_INTERIOR_EXPLICIT):
        for b in range(CONTRA_MAX_INTERIOR_EXPLICIT):
            ex[a, b] = -0.5 + 1.0 * _hash01(66, min(a, b), max(a, b))
    v = view(t, "bulge_scores_0x1")
    for x in range(4):
        v[x] = -0.3 + 0.5 * _hash01(67, x)
    v = view(t, "interior_scores_1x1")
    for x in range(4):
        for y in range(4):
            v[x, y] = -0.6 + 1.4 * _hash01(68, min(x, y), max(x, y))
    t.multibranch_score_base = -1.199055076
    t.multibranch_score_basepair = -0.9253883752
    t.multibranch_score_unpair = -0.1983300391
    t.external_score_basepair = -0.0009674111431
    t.external_score_unpair = -0.00972883093
    accumulate(t, lib)
    return t


def accumulate(t: ContraTables, lib=None) -> None:
    """FoldScoreSets::accumulate (src/mccaskill_algo.rs:60-86): sequential f32 prefix sums.

    Uses the C ABI helper when a loaded library is given, else an equivalent numpy f32 cumsum
    (np.cumsum on float32 accumulates sequentially in float32 for 1-D arrays of this size)."""
    if lib is not None:
        lib.rna_contra_tables_accumulate(C.byref(t))
        return
    for name in ("hairpin_scores_len", "bulge_scores_len", "interior_scores_len",
                 "interior_scores_symmetric", "interior_scores_asymmetric"):
        src = view(t, name)
        dst = view(t, name + "_cumulative")
        s = np.float32(0.0)
        for n in range(src.shape[0]):
            s = np.float32(s + src[n])
            dst[n] = s


# ----------------------------------------------------------------------------------------------
# Fuzz tables
# ----------------------------------------------------------------------------------------------
def random_turner_tables(seed: int, special: bool = True) -> TurnerTables:
    rng = np.random.default_rng(seed)
    t = TurnerTables()
    t.max_2loop_len = int(rng.integers(3, 31))
    t.min_span_hairpin_close = int(rng.integers(4, 7))
    t.min_hairpin_len = t.min_span_hairpin_close - 2
    t.max_hairpin_len_extrapolation = int(rng.integers(6, 12))
    t.min_hairpin_len_extrapolation = t.max_hairpin_len_extrapolation + 1
    for name in ("coeff_hairpin_len_extrapolation", "helix_augu_end_penalty", "ninio_coeff", "ninio_max",
                 "init_multibranch_base", "coeff_num_branches"):
        setattr(t, name, float(np.float32(rng.uniform(-3.0, 1.0))))
    t.ninio_max = float(np.float32(rng.uniform(-5.0, -1.0)))
    t.ninio_coeff = float(np.float32(rng.uniform(-1.5, -0.2)))
    for name, _ in TurnerTables._fields_:
        fld = getattr(t, name)
        if isinstance(fld, C.Array) and name != "hairpin_scores_special":
            v = view(t, name)
            v[...] = rng.uniform(-4.0, 3.0, size=v.shape).astype(np.float32)
    n = 0
    if special:
        for ln in (t.min_span_hairpin_close, t.min_span_hairpin_close + 1, 8):
            for _ in range(6):
                ent = t.hairpin_scores_special[n]
                ent.len = ln
                closing = sorted(CANONICAL)[int(rng.integers(0, 6))]
                ent.seq[0] = closing[0]
                ent.seq[ln - 1] = closing[1]
                for p in range(1, ln - 1):
                    ent.seq[p] = int(rng.integers(0, 4))
                ent.score = float(np.float32(rng.uniform(-8.0, -2.0)))
                n += 1
    t.num_special_hairpins = n
    return t


def random_contra_tables(seed: int) -> ContraTables:
    rng = np.random.default_rng(seed)
    t = ContraTables()
    t.max_loop_len = CONTRA_MAX_LOOP_LEN
    t.min_span_hairpin_close = int(rng.integers(4, 7))
    t.max_interior_explicit = int(rng.integers(1, 5))
    for name, _ in ContraTables._fields_:
        fld = getattr(t, name)
        if isinstance(fld, C.Array) and not name.endswith("_cumulative"):
            v = view(t, name)
            v[...] = rng.uniform(-1.5, 1.5, size=v.shape).astype(np.float32)
    for name in ("multibranch_score_base", "multibranch_score_basepair", "multibranch_score_unpair",
                 "external_score_basepair", "external_score_unpair"):
        setattr(t, name, float(np.float32(rng.uniform(-1.5, 0.3))))
    accumulate(t)
    return t


def random_align_tables(seed: int) -> AlignTables:
    rng = np.random.default_rng(seed)
    t = AlignTables()
    for name in ("match2match_score", "match2insert_score", "insert_extend_score", "insert_switch_score",
                 "init_match_score", "init_insert_score"):
        setattr(t, name, float(np.float32(rng.uniform(-3.0, 3.0))))
    view(t, "insert_scores")[...] = rng.uniform(-1.0, 1.0, size=4).astype(np.float32)
    m = rng.uniform(-1.0, 1.0, size=(4, 4)).astype(np.float32)
    view(t, "match_scores")[...] = (m + m.T) / 2
    return t


# ----------------------------------------------------------------------------------------------
# Table blob files: the raw bytes of the C structs behind a 16-byte header
#   "RNATBL01" | u32 kind (1 Turner, 2 CONTRAfold, 3 align) | u32 sizeof(struct)
# File names `turner2004.tbl` / `contrafold_v202.tbl` are RESERVED for the genuine rna-ss-params values (written by
# tools/ref_dump on a machine with cargo); `python -m rna_algos_b200.tables dump DIR` writes the stand-ins as
# `standin_turner.tbl` / `standin_contrafold.tbl`.
# ----------------------------------------------------------------------------------------------
GENUINE_TURNER, GENUINE_CONTRA = "turner2004.tbl", "contrafold_v202.tbl"
STANDIN_TURNER, STANDIN_CONTRA = "standin_turner.tbl", "standin_contrafold.tbl"
_KIND_STRUCT = {1: TurnerTables, 2: ContraTables, 3: AlignTables}


def dump_table_file(path: str, kind: int, struct) -> None:
    import struct as S
    raw = bytes(struct)
    with open(path, "wb") as f:
        f.write(b"RNATBL01" + S.pack("<II", kind, len(raw)) + raw)


def load_table_file(path: str, kind: int):
    import struct as S
    with open(path, "rb") as f:
        head = f.read(16)
        raw = f.read()
    if len(head) != 16 or head[:8] != b"RNATBL01":
        raise ValueError(f"{path}: not an RNATBL01 table blob")
    k, size = S.unpack("<II", head[8:])
    cls = _KIND_STRUCT[kind]
    if k != kind or size != C.sizeof(cls) or len(raw) != size:
        raise ValueError(f"{path}: kind/size mismatch (kind {k}, {size} bytes; expected kind {kind}, {C.sizeof(cls)} bytes)")
    return cls.from_buffer_copy(raw)


def load_genuine_tables(directory: str):
    """(TurnerTables, ContraTables) from the reserved file names in `directory`."""
    import os
    return (load_table_file(os.path.join(directory, GENUINE_TURNER), 1),
            load_table_file(os.path.join(directory, GENUINE_CONTRA), 2))


def dump_standin_tables(out_dir: str) -> None:
    import os
    os.makedirs(out_dir, exist_ok=True)
    dump_table_file(os.path.join(out_dir, STANDIN_TURNER), 1, standin_turner_tables())
    dump_table_file(os.path.join(out_dir, STANDIN_CONTRA), 2, standin_contra_tables())


if __name__ == "__main__":
    import sys
    if len(sys.argv) == 3 and sys.argv[1] == "dump":
        dump_standin_tables(sys.argv[2])
    else:
        raise SystemExit("usage: python -m rna_algos_b200.tables dump DIR   (writes the STAND-IN blobs)")
