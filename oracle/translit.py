"""translit.py — a SECOND, independent restatement of the reference's hot path.  TEST INFRASTRUCTURE ONLY.

oracle.c restates heartsh/rna-algos 0.1.37 in C with dense arrays.  This file was written separately, straight from
the Rust text, as a literal line-by-line transliteration that keeps the reference's own data structures: sparse
`HashMap`s become dicts keyed by the same tuples (`sums_close`, `sums_accessible`, `basepair_probs`, the four
`FoldScores` memo maps incl. the 4-D `twoloop_scores`), dense `Vec<Vec<f32>>` become lists of lists, a missing key
raises KeyError exactly where the Rust would panic, and every arithmetic operation is one IEEE f32 operation
(numpy float32 scalars; no fused multiply-add exists in this interpreter).  `f32::ln` / `f32::exp` go to the C
library's logf / expf like Rust's std does.  tests/test_second_restatement.py asserts that the two restatements
agree BIT FOR BIT on the bundled tRNAs and on random sequences with random tables: with the reference itself
unbuildable here (no cargo, table crate not vendored), two independent readings of the same text agreeing is the
evidence available.  Pure-Python loops: small inputs only.

Every function cites the reference lines it follows (paths relative to the reference tree).
"""
from __future__ import annotations

import ctypes
import ctypes.util

import numpy as np

f32 = np.float32
NEG_INFINITY = f32(-np.inf)
ZERO = f32(0.0)

_libm = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
_libm.logf.restype = ctypes.c_float
_libm.logf.argtypes = [ctypes.c_float]
_libm.expf.restype = ctypes.c_float
_libm.expf.argtypes = [ctypes.c_float]

A, C, G, U = 0, 1, 2, 3
AU, CG, GC, GU, UA, UG = (A, U), (C, G), (G, C), (G, U), (U, A), (U, G)
PSEUDO_BASE = U + 1                       # src/utils.rs:122
LOGSUMEXP_THRESHOLD_UPPER = f32(11.862479)  # src/utils.rs:121


def _is_finite(x) -> bool:
    return bool(np.isfinite(x))


# ---------------------------------------------------------------------------------------------------------
# src/utils.rs:579-655
# ---------------------------------------------------------------------------------------------------------
def ln_exp_1p(x):
    """src/utils.rs:602-627 (the literals are the reference's, digit for digit)"""
    if x < f32(3.37925):
        if x < f32(1.6320158):
            if x < f32(0.66153675):
                return ((f32(-0.0065591595) * x + f32(0.12764427)) * x + f32(0.49965546)) * x + f32(0.6931542)
            else:
                return ((f32(-0.015515756) * x + f32(0.14467756)) * x + f32(0.48829398)) * x + f32(0.6958093)
        elif x < f32(2.4912589):
            return ((f32(-0.012890925) * x + f32(0.13010283)) * x + f32(0.51503986)) * x + f32(0.6795586)
        else:
            return ((f32(-0.0072142647) * x + f32(0.087754086)) * x + f32(0.6208708)) * x + f32(0.5909676)
    elif x < f32(5.789071):
        if x < f32(4.426169):
            return ((f32(-0.0031455354) * x + f32(0.046722945)) * x + f32(0.7592532)) * x + f32(0.43487945)
        else:
            return ((f32(-0.0010110698) * x + f32(0.018594341)) * x + f32(0.88317305)) * x + f32(0.25236955)
    elif x < f32(7.8162727):
        return ((f32(-0.000196278) * x + f32(0.0046084408)) * x + f32(0.9634432)) * x + f32(0.09831489)
    else:
        return ((f32(-0.0000113994) * x + f32(0.0003734731)) * x + f32(0.9959107)) * x + f32(0.0149855051)


def logsumexp(sum_, x):
    """src/utils.rs:580-596; returns the new sum (the Rust mutates `*sum`)."""
    if not _is_finite(x):
        return sum_
    if not _is_finite(sum_):
        return x
    y = min(sum_, x)
    z = max(sum_, x) - y
    return y + (z if z >= LOGSUMEXP_THRESHOLD_UPPER else ln_exp_1p(z))


def expf(x):
    """src/utils.rs:630-655"""
    if x < f32(-2.4915035):
        if x < f32(-5.8622823):
            if x < f32(-9.91152):
                return ZERO
            else:
                return ((f32(0.0000803850) * x + f32(0.002162743)) * x + f32(0.019470856)) * x + f32(0.058808003)
        elif x < f32(-3.839663):
            return ((f32(0.0013889414) * x + f32(0.024467647)) * x + f32(0.14712906)) * x + f32(0.30427578)
        else:
            return ((f32(0.0072335607) * x + f32(0.09060027)) * x + f32(0.39831114)) * x + f32(0.62459594)
    elif x < f32(-0.6725053):
        if x < f32(-1.4805375):
            return ((f32(0.023241036) * x + f32(0.2085646)) * x + f32(0.6906368)) * x + f32(0.86823225)
        else:
            return ((f32(0.057378277) * x + f32(0.35802585)) * x + f32(0.9121133)) * x + f32(0.9793092)
    elif x < ZERO:
        return ((f32(0.119917594) * x + f32(0.48156682)) * x + f32(0.9975992)) * x + f32(0.9999505)
    else:
        return f32(_libm.expf(float(x)))


# ---------------------------------------------------------------------------------------------------------
# rna-ss-params symbols, read from the run-time table blobs (include/rna_algos_b200.h)
# ---------------------------------------------------------------------------------------------------------
def _arr(struct, name):
    fld = getattr(struct, name)
    return np.ctypeslib.as_array(fld).astype(np.float32)


class Turner:
    """The compiled_scores_turner / utils constants the Turner path reads (src/utils.rs:8-10 glob imports)."""

    def __init__(self, t):
        self.MAX_2LOOP_LEN = int(t.max_2loop_len)
        self.MIN_SPAN_HAIRPIN_CLOSE = int(t.min_span_hairpin_close)
        self.MIN_HAIRPIN_LEN = int(t.min_hairpin_len)
        self.MAX_HAIRPIN_LEN_EXTRAPOLATION = int(t.max_hairpin_len_extrapolation)
        self.MIN_HAIRPIN_LEN_EXTRAPOLATION = int(t.min_hairpin_len_extrapolation)
        self.COEFF_HAIRPIN_LEN_EXTRAPOLATION = f32(t.coeff_hairpin_len_extrapolation)
        self.HELIX_AUGU_END_PENALTY = f32(t.helix_augu_end_penalty)
        self.NINIO_COEFF = f32(t.ninio_coeff)
        self.NINIO_MAX = f32(t.ninio_max)
        self.INIT_MULTIBRANCH_BASE = f32(t.init_multibranch_base)
        self.COEFF_NUM_BRANCHES = f32(t.coeff_num_branches)
        self.HAIRPIN_SCORES_INIT = _arr(t, "hairpin_scores_init")
        self.BULGE_SCORES_INIT = _arr(t, "bulge_scores_init")
        self.INTERIOR_SCORES_INIT = _arr(t, "interior_scores_init")
        self.STACK_SCORES = _arr(t, "stack_scores")
        self.TERMINAL_MISMATCH_SCORES_HAIRPIN = _arr(t, "terminal_mismatch_scores_hairpin")
        self.TERMINAL_MISMATCH_SCORES_1XMANY = _arr(t, "terminal_mismatch_scores_1xmany")
        self.TERMINAL_MISMATCH_SCORES_2X3 = _arr(t, "terminal_mismatch_scores_2x3")
        self.TERMINAL_MISMATCH_SCORES_INTERIOR = _arr(t, "terminal_mismatch_scores_interior")
        self.TERMINAL_MISMATCH_SCORES_MULTIBRANCH = _arr(t, "terminal_mismatch_scores_multibranch")
        self.DANGLING_SCORES_5PRIME = _arr(t, "dangling_scores_5prime")
        self.DANGLING_SCORES_3PRIME = _arr(t, "dangling_scores_3prime")
        self.INTERIOR_SCORES_1X1 = _arr(t, "interior_scores_1x1")
        self.INTERIOR_SCORES_1X2 = _arr(t, "interior_scores_1x2")
        self.INTERIOR_SCORES_2X2 = _arr(t, "interior_scores_2x2")
        self.HAIRPIN_SCORES_SPECIAL = []
        for x in range(int(t.num_special_hairpins)):
            e = t.hairpin_scores_special[x]
            self.HAIRPIN_SCORES_SPECIAL.append((tuple(int(b) for b in e.seq[: e.len]), f32(e.score)))


class FoldScoreSets:
    """src/utils.rs:91-119: the fields of the CONTRAfold parameter set, as left by ::new(0.).transfer()."""

    def __init__(self, c):
        self.MAX_LOOP_LEN = int(c.max_loop_len)
        self.MIN_SPAN_HAIRPIN_CLOSE = int(c.min_span_hairpin_close)
        self.MAX_INTERIOR_EXPLICIT = int(c.max_interior_explicit)
        for name in ("hairpin_scores_len_cumulative", "bulge_scores_len_cumulative", "interior_scores_len_cumulative",
                     "interior_scores_symmetric_cumulative", "interior_scores_asymmetric_cumulative", "stack_scores",
                     "terminal_mismatch_scores", "dangling_scores_left", "dangling_scores_right", "helix_close_scores",
                     "basepair_scores", "interior_scores_explicit", "bulge_scores_0x1", "interior_scores_1x1"):
            setattr(self, name, _arr(c, name))
        for name in ("multibranch_score_base", "multibranch_score_basepair", "multibranch_score_unpair",
                     "external_score_basepair", "external_score_unpair"):
            setattr(self, name, f32(getattr(c, name)))


# ---------------------------------------------------------------------------------------------------------
# Turner scorers, src/utils.rs:162-411
# ---------------------------------------------------------------------------------------------------------
def has_canonical_basepair(x) -> bool:
    return x in (AU, CG, GC, GU, UA, UG)           # src/utils.rs:162-164


def matches_augu(x) -> bool:
    return x == AU or x == UA or x == GU or x == UG  # src/utils.rs:558-560


def invert_basepair(x):
    return (x[1], x[0])


def get_abs_diff(x, y):
    return max(x, y) - min(x, y)


def get_special_hairpin_score(P: Turner, seq):
    for x in P.HAIRPIN_SCORES_SPECIAL:             # src/utils.rs:198-205
        if x[0] == tuple(seq):
            return x[1]
    return NEG_INFINITY


def get_hairpin_score(P: Turner, seq, pos_pair_close):
    """src/utils.rs:166-196"""
    hairpin = seq[pos_pair_close[0]: pos_pair_close[1] + 1]
    special_hairpin_score = get_special_hairpin_score(P, hairpin)
    if special_hairpin_score > NEG_INFINITY:
        return special_hairpin_score
    hairpin_len = pos_pair_close[1] - pos_pair_close[0] - 1
    basepair_close = (seq[pos_pair_close[0]], seq[pos_pair_close[1]])
    if hairpin_len == P.MIN_HAIRPIN_LEN:
        hairpin_score = P.HAIRPIN_SCORES_INIT[hairpin_len]
    else:
        terminal_mismatch = (seq[pos_pair_close[0] + 1], seq[pos_pair_close[1] - 1])
        if hairpin_len <= P.MAX_HAIRPIN_LEN_EXTRAPOLATION:
            hairpin_score_init = P.HAIRPIN_SCORES_INIT[hairpin_len]
        else:
            ratio = f32(hairpin_len) / f32(P.MIN_HAIRPIN_LEN_EXTRAPOLATION - 1)
            hairpin_score_init = (P.HAIRPIN_SCORES_INIT[P.MIN_HAIRPIN_LEN_EXTRAPOLATION - 1]
                                  + P.COEFF_HAIRPIN_LEN_EXTRAPOLATION * f32(_libm.logf(float(ratio))))
        hairpin_score = hairpin_score_init + P.TERMINAL_MISMATCH_SCORES_HAIRPIN[basepair_close[0]][basepair_close[1]][
            terminal_mismatch[0]][terminal_mismatch[1]]
    return hairpin_score + (P.HELIX_AUGU_END_PENALTY if matches_augu(basepair_close) else ZERO)


def get_stack_score(P, seq, pc, pa):
    bc = (seq[pc[0]], seq[pc[1]])                   # src/utils.rs:224-232
    ba = (seq[pa[0]], seq[pa[1]])
    return P.STACK_SCORES[bc[0]][bc[1]][ba[0]][ba[1]]


def get_bulge_score(P, seq, pc, pa):
    """src/utils.rs:234-258"""
    bulge_len = pa[0] - pc[0] + pc[1] - pa[1] - 2
    if bulge_len == 1:
        return P.BULGE_SCORES_INIT[bulge_len] + get_stack_score(P, seq, pc, pa)
    bc = (seq[pc[0]], seq[pc[1]])
    ba = (seq[pa[0]], seq[pa[1]])
    return (P.BULGE_SCORES_INIT[bulge_len]
            + (P.HELIX_AUGU_END_PENALTY if matches_augu(bc) else ZERO)
            + (P.HELIX_AUGU_END_PENALTY if matches_augu(ba) else ZERO))


def get_interior_mismatch_score(P, seq, pc, pa, loop_len_pair):
    """src/utils.rs:331-366"""
    bc = (seq[pc[0]], seq[pc[1]])
    ba = (seq[pa[1]], seq[pa[0]])
    tm = ((seq[pc[0] + 1], seq[pc[1] - 1]), (seq[pa[1] + 1], seq[pa[0] - 1]))
    if loop_len_pair[0] == 1 or loop_len_pair[1] == 1:
        tab = P.TERMINAL_MISMATCH_SCORES_1XMANY
    elif loop_len_pair in ((2, 3), (3, 2)):
        tab = P.TERMINAL_MISMATCH_SCORES_2X3
    else:
        tab = P.TERMINAL_MISMATCH_SCORES_INTERIOR
    return tab[bc[0]][bc[1]][tm[0][0]][tm[0][1]] + tab[ba[0]][ba[1]][tm[1][0]][tm[1][1]]


def get_interior_score(P, seq, pc, pa):
    """src/utils.rs:260-321"""
    bc = (seq[pc[0]], seq[pc[1]])
    ba = (seq[pa[0]], seq[pa[1]])
    loop_len_pair = (pa[0] - pc[0] - 1, pc[1] - pa[1] - 1)
    interior_len = loop_len_pair[0] + loop_len_pair[1]
    if loop_len_pair == (1, 1):
        interior = (seq[pc[0] + 1], seq[pc[1] - 1])
        return P.INTERIOR_SCORES_1X1[bc[0]][bc[1]][interior[0]][interior[1]][ba[0]][ba[1]]
    if loop_len_pair == (1, 2):
        interior = ((seq[pc[0] + 1], seq[pc[1] - 1]), seq[pc[1] - 2])
        return P.INTERIOR_SCORES_1X2[bc[0]][bc[1]][interior[0][0]][interior[0][1]][interior[1]][ba[0]][ba[1]]
    if loop_len_pair == (2, 1):
        interior = ((seq[pc[1] - 1], seq[pc[0] + 2]), seq[pc[0] + 1])
        bai = invert_basepair(ba)
        bci = invert_basepair(bc)
        return P.INTERIOR_SCORES_1X2[bai[0]][bai[1]][interior[0][0]][interior[0][1]][interior[1]][bci[0]][bci[1]]
    if loop_len_pair == (2, 2):
        interior = ((seq[pc[0] + 1], seq[pc[1] - 1]), (seq[pc[0] + 2], seq[pc[1] - 2]))
        return P.INTERIOR_SCORES_2X2[bc[0]][bc[1]][interior[0][0]][interior[0][1]][interior[1][0]][interior[1][1]][
            ba[0]][ba[1]]
    ninio = max(P.NINIO_COEFF * f32(get_abs_diff(loop_len_pair[0], loop_len_pair[1])), P.NINIO_MAX)
    return (P.INTERIOR_SCORES_INIT[interior_len]
            + ninio
            + get_interior_mismatch_score(P, seq, pc, pa, loop_len_pair)
            + (P.HELIX_AUGU_END_PENALTY if matches_augu(bc) else ZERO)
            + (P.HELIX_AUGU_END_PENALTY if matches_augu(ba) else ZERO))


def get_2loop_score(P, seq, pc, pa):
    """src/utils.rs:207-222"""
    if pc[0] + 1 == pa[0] and pc[1] - 1 == pa[1]:
        return get_stack_score(P, seq, pc, pa)
    elif pc[0] + 1 == pa[0] or pc[1] - 1 == pa[1]:
        return get_bulge_score(P, seq, pc, pa)
    return get_interior_score(P, seq, pc, pa)


def get_multibranch_close_score(P, seq, pc):
    """src/utils.rs:368-382"""
    bc = (seq[pc[0]], seq[pc[1]])
    bci = invert_basepair(bc)
    bsi = invert_basepair((seq[pc[0] + 1], seq[pc[1] - 1]))
    tms = P.TERMINAL_MISMATCH_SCORES_MULTIBRANCH[bci[0]][bci[1]][bsi[0]][bsi[1]]
    return P.INIT_MULTIBRANCH_BASE + tms + (P.HELIX_AUGU_END_PENALTY if matches_augu(bc) else ZERO)


def get_accessible_score(P, seq, pa, uses_sentinel_bases):
    """src/utils.rs:384-411"""
    seq_len = len(seq)
    end_5prime = 1 if uses_sentinel_bases else 0
    end_3prime = seq_len - (2 if uses_sentinel_bases else 1)
    ba = (seq[pa[0]], seq[pa[1]])
    if pa[0] > end_5prime and pa[1] < end_3prime:
        score = P.TERMINAL_MISMATCH_SCORES_MULTIBRANCH[ba[0]][ba[1]][seq[pa[0] - 1]][seq[pa[1] + 1]]
    elif pa[0] > end_5prime:
        score = P.DANGLING_SCORES_5PRIME[ba[0]][ba[1]][seq[pa[0] - 1]]
    elif pa[1] < end_3prime:
        score = P.DANGLING_SCORES_3PRIME[ba[0]][ba[1]][seq[pa[1] + 1]]
    else:
        score = ZERO
    return score + (P.HELIX_AUGU_END_PENALTY if matches_augu(ba) else ZERO)


# ---------------------------------------------------------------------------------------------------------
# CONTRAfold scorers, src/utils.rs:413-556
# ---------------------------------------------------------------------------------------------------------
def get_helix_close_score(x, F):
    return F.helix_close_scores[x[0]][x[1]]


def get_terminal_mismatch_score(x, y, F):
    return F.terminal_mismatch_scores[x[0]][x[1]][y[0]][y[1]]


def get_junction_score_single(x, y, F):
    a = (x[y[0]], x[y[1]])                          # src/utils.rs:545-548
    return get_helix_close_score(a, F) + get_terminal_mismatch_score(a, (x[y[0] + 1], x[y[1] - 1]), F)


def get_junction_score(seq, pos_pair, uses_sentinel_bases, F):
    """src/utils.rs:522-543"""
    seq_len = len(seq)
    basepair = (seq[pos_pair[0]], seq[pos_pair[1]])
    end_5prime = 1 if uses_sentinel_bases else 0
    end_3prime = seq_len - (2 if uses_sentinel_bases else 1)
    return (get_helix_close_score(basepair, F)
            + (F.dangling_scores_left[basepair[0]][basepair[1]][seq[pos_pair[0] + 1]] if pos_pair[0] < end_3prime else ZERO)
            + (F.dangling_scores_right[basepair[0]][basepair[1]][seq[pos_pair[1] - 1]] if pos_pair[1] > end_5prime else ZERO))


def get_hairpin_score_contra(seq, pc, F):
    hairpin_len = pc[1] - pc[0] - 1                 # src/utils.rs:413-421
    return F.hairpin_scores_len_cumulative[min(hairpin_len, F.MAX_LOOP_LEN)] + get_junction_score_single(seq, pc, F)


def get_stack_score_contra(seq, pc, pa, F):
    bc = (seq[pc[0]], seq[pc[1]])                   # src/utils.rs:444-454
    ba = (seq[pa[0]], seq[pa[1]])
    return F.stack_scores[bc[0]][bc[1]][ba[0]][ba[1]]


def get_bulge_score_contra(seq, pc, pa, F):
    """src/utils.rs:456-481"""
    bulge_len = pa[0] - pc[0] + pc[1] - pa[1] - 2
    if bulge_len == 1:
        score = F.bulge_scores_0x1[seq[pc[0] + 1] if pa[0] - pc[0] - 1 == 1 else seq[pc[1] - 1]]
    else:
        score = ZERO
    return (score
            + F.bulge_scores_len_cumulative[bulge_len - 1]
            + get_junction_score_single(seq, pc, F)
            + get_junction_score_single(seq, (pa[1], pa[0]), F))


def get_interior_score_contra(seq, pc, pa, F):
    """src/utils.rs:483-520"""
    loop_len_pair = (pa[0] - pc[0] - 1, pc[1] - pa[1] - 1)
    interior_len = loop_len_pair[0] + loop_len_pair[1]
    if loop_len_pair[0] == loop_len_pair[1]:
        score_1x1 = F.interior_scores_1x1[seq[pc[0] + 1]][seq[pc[1] - 1]] if interior_len == 2 else ZERO
        score = score_1x1 + F.interior_scores_symmetric_cumulative[loop_len_pair[0] - 1]
    else:
        score = F.interior_scores_asymmetric_cumulative[get_abs_diff(loop_len_pair[0], loop_len_pair[1]) - 1]
    if loop_len_pair[0] <= F.MAX_INTERIOR_EXPLICIT and loop_len_pair[1] <= F.MAX_INTERIOR_EXPLICIT:
        score_explicit = F.interior_scores_explicit[loop_len_pair[0] - 1][loop_len_pair[1] - 1]
    else:
        score_explicit = ZERO
    return (score
            + score_explicit
            + F.interior_scores_len_cumulative[interior_len - 2]
            + get_junction_score_single(seq, pc, F)
            + get_junction_score_single(seq, (pa[1], pa[0]), F))


def get_2loop_score_contra(seq, pc, pa, F):
    """src/utils.rs:423-442"""
    ba = (seq[pa[0]], seq[pa[1]])
    if pc[0] + 1 == pa[0] and pc[1] - 1 == pa[1]:
        score = get_stack_score_contra(seq, pc, pa, F)
    elif pc[0] + 1 == pa[0] or pc[1] - 1 == pa[1]:
        score = get_bulge_score_contra(seq, pc, pa, F)
    else:
        score = get_interior_score_contra(seq, pc, pa, F)
    return score + F.basepair_scores[ba[0]][ba[1]]


# ---------------------------------------------------------------------------------------------------------
# src/mccaskill_algo.rs
# ---------------------------------------------------------------------------------------------------------
class FoldSums:
    """src/mccaskill_algo.rs:3-11, 213-226"""

    def __init__(self, seq_len):
        neg = lambda: [[NEG_INFINITY] * seq_len for _ in range(seq_len)]
        self.sums_external = [[ZERO] * seq_len for _ in range(seq_len)]
        self.sums_rightmost_basepairs_external = neg()
        self.sums_rightmost_basepairs_multibranch = neg()
        self.sums_close = {}
        self.sums_accessible = {}
        self.sums_multibranch = neg()
        self.sums_1ormore_basepairs = neg()


class FoldScores:
    """src/mccaskill_algo.rs:13-22, 234-245"""

    def __init__(self):
        self.hairpin_scores = {}
        self.twoloop_scores = {}
        self.multibranch_close_scores = {}
        self.accessible_scores = {}


def get_fold_sums(seq, fold_scores: FoldScores, P: Turner) -> FoldSums:
    """src/mccaskill_algo.rs:282-378"""
    seq_len = len(seq)
    uses_sentinel_bases = False
    fs = FoldSums(seq_len)
    for subseq_len in range(P.MIN_SPAN_HAIRPIN_CLOSE, seq_len + 1):
        for i in range(0, seq_len - subseq_len + 1):
            j = i + subseq_len - 1
            pos_pair_close = (i, j)
            basepair_close = (seq[i], seq[j])
            sum_ = NEG_INFINITY
            if j - i + 1 >= P.MIN_SPAN_HAIRPIN_CLOSE and has_canonical_basepair(basepair_close):
                hairpin_score = get_hairpin_score(P, seq, pos_pair_close)
                fold_scores.hairpin_scores[pos_pair_close] = hairpin_score
                sum_ = logsumexp(sum_, hairpin_score)
                for k in range(i + 1, j - 1):
                    if k - i - 1 > P.MAX_2LOOP_LEN:
                        break
                    for l in reversed(range(k + 1, j)):
                        if j - l - 1 + k - i - 1 > P.MAX_2LOOP_LEN:
                            break
                        pos_pair_accessible = (k, l)
                        x = fs.sums_close.get(pos_pair_accessible)
                        if x is not None:
                            y = get_2loop_score(P, seq, pos_pair_close, pos_pair_accessible)
                            fold_scores.twoloop_scores[(i, j, k, l)] = y
                            y = x + y
                            sum_ = logsumexp(sum_, y)
                multibranch_close_score = get_multibranch_close_score(P, seq, pos_pair_close)
                sum_ = logsumexp(sum_, fs.sums_multibranch[i + 1][j - 1] + multibranch_close_score)
                accessible_score = get_accessible_score(P, seq, pos_pair_close, uses_sentinel_bases)
                if sum_ > NEG_INFINITY:
                    fold_scores.multibranch_close_scores[pos_pair_close] = multibranch_close_score
                    fold_scores.accessible_scores[pos_pair_close] = accessible_score
                    fs.sums_close[pos_pair_close] = sum_
                    fs.sums_accessible[pos_pair_close] = sum_ + accessible_score
            sum_ = NEG_INFINITY
            for k in range(i + 1, j + 1):
                x = fs.sums_accessible.get((i, k))
                if x is not None:
                    sum_ = logsumexp(sum_, x)
            fs.sums_rightmost_basepairs_external[i][j] = sum_
            sum_ = ZERO
            for k in range(i, j):
                x = fs.sums_rightmost_basepairs_external[k][j]
                y = ZERO if (i == 0 and k == 0) else fs.sums_external[i][k - 1]
                y = x + y
                sum_ = logsumexp(sum_, y)
            fs.sums_external[i][j] = sum_
            sum_ = fs.sums_rightmost_basepairs_external[i][j] + P.COEFF_NUM_BRANCHES
            sum2 = NEG_INFINITY
            for k in range(i + 1, j):
                x = fs.sums_rightmost_basepairs_external[k][j] + P.COEFF_NUM_BRANCHES
                sum_ = logsumexp(sum_, x)
                y = fs.sums_1ormore_basepairs[i][k - 1] + x
                sum2 = logsumexp(sum2, y)
            fs.sums_multibranch[i][j] = sum2
            sum_ = logsumexp(sum_, sum2)
            fs.sums_1ormore_basepairs[i][j] = sum_
    return fs


def get_fold_sums_contra(seq, fold_scores: FoldScores, allows_short_hairpins: bool, F: FoldScoreSets) -> FoldSums:
    """src/mccaskill_algo.rs:380-516"""
    seq_len = len(seq)
    uses_sentinel_bases = False
    fs = FoldSums(seq_len)
    for subseq_len in range(1, seq_len + 1):
        for i in range(0, seq_len - subseq_len + 1):
            j = i + subseq_len - 1
            pos_pair_close = (i, j)
            basepair_close = (seq[i], seq[j])
            sum_ = NEG_INFINITY
            if has_canonical_basepair(basepair_close) and (allows_short_hairpins or j - i + 1 >= F.MIN_SPAN_HAIRPIN_CLOSE):
                if j - i - 1 <= F.MAX_LOOP_LEN:
                    hairpin_score = get_hairpin_score_contra(seq, pos_pair_close, F)
                    fold_scores.hairpin_scores[pos_pair_close] = hairpin_score
                    sum_ = logsumexp(sum_, hairpin_score)
                for k in range(i + 1, j - 1):
                    if k - i - 1 > F.MAX_LOOP_LEN:
                        break
                    for l in reversed(range(k + 1, j)):
                        if j - l - 1 + k - i - 1 > F.MAX_LOOP_LEN:
                            break
                        pos_pair_accessible = (k, l)
                        x = fs.sums_close.get(pos_pair_accessible)
                        if x is not None:
                            y = get_2loop_score_contra(seq, pos_pair_close, pos_pair_accessible, F)
                            fold_scores.twoloop_scores[(i, j, k, l)] = y
                            y = x + y
                            sum_ = logsumexp(sum_, y)
                multibranch_close_score = (F.multibranch_score_base + F.multibranch_score_basepair
                                           + get_junction_score(seq, pos_pair_close, uses_sentinel_bases, F))
                sum_ = logsumexp(sum_, fs.sums_multibranch[i + 1][j - 1] + multibranch_close_score)
                accessible_score = (get_junction_score(seq, (pos_pair_close[1], pos_pair_close[0]), uses_sentinel_bases, F)
                                    + F.basepair_scores[basepair_close[0]][basepair_close[1]])
                if sum_ > NEG_INFINITY:
                    fold_scores.multibranch_close_scores[pos_pair_close] = multibranch_close_score
                    fold_scores.accessible_scores[pos_pair_close] = accessible_score
                    fs.sums_close[pos_pair_close] = sum_
                    fs.sums_accessible[pos_pair_close] = sum_ + accessible_score
            sum_ = NEG_INFINITY
            sum2 = sum_
            for k in range(i + 1, j + 1):
                x = fs.sums_accessible.get((i, k))
                if x is not None:
                    sum_ = logsumexp(sum_, x + F.external_score_basepair + F.external_score_unpair * f32(j - k))
                    sum2 = logsumexp(sum2, x + F.multibranch_score_basepair + F.multibranch_score_unpair * f32(j - k))
            fs.sums_rightmost_basepairs_external[i][j] = sum_
            fs.sums_rightmost_basepairs_multibranch[i][j] = sum2
            sum_ = F.external_score_unpair * f32(subseq_len)
            for k in range(i, j):
                x = fs.sums_rightmost_basepairs_external[k][j]
                y = ZERO if (i == 0 and k == 0) else fs.sums_external[i][k - 1]
                y = x + y
                sum_ = logsumexp(sum_, y)
            fs.sums_external[i][j] = sum_
            sum_ = fs.sums_rightmost_basepairs_multibranch[i][j]
            sum2 = NEG_INFINITY
            for k in range(i + 1, j):
                x = fs.sums_rightmost_basepairs_multibranch[k][j]
                sum_ = logsumexp(sum_, x + F.multibranch_score_unpair * f32(k - i))
                x = fs.sums_1ormore_basepairs[i][k - 1] + x
                sum2 = logsumexp(sum2, x)
            fs.sums_multibranch[i][j] = sum2
            sum_ = logsumexp(sum_, sum2)
            fs.sums_1ormore_basepairs[i][j] = sum_
    return fs


def get_basepair_probs(fs: FoldSums, seq_len: int, fold_scores: FoldScores, P: Turner):
    """src/mccaskill_algo.rs:518-610"""
    global_sum = fs.sums_external[0][seq_len - 1]
    basepair_probs = {}
    probs_multibranch = [[NEG_INFINITY] * seq_len for _ in range(seq_len)]
    probs_multibranch2 = [[NEG_INFINITY] * seq_len for _ in range(seq_len)]
    for subseq_len in reversed(range(P.MIN_SPAN_HAIRPIN_CLOSE, seq_len + 1)):
        for i in range(0, seq_len - subseq_len + 1):
            j = i + subseq_len - 1
            sum_ = NEG_INFINITY
            sum2 = sum_
            for k in range(j + 1, seq_len):
                pos_pair_close = (i, k)
                x = fs.sums_close.get(pos_pair_close)
                if x is not None:
                    basepair_prob = basepair_probs[pos_pair_close]
                    multibranch_close_score = fold_scores.multibranch_close_scores[pos_pair_close]
                    x = basepair_prob + multibranch_close_score - x
                    sum_ = logsumexp(sum_, x + fs.sums_1ormore_basepairs[j + 1][k - 1])
                    sum2 = logsumexp(sum2, x)
            probs_multibranch[i][j] = sum_
            probs_multibranch2[i][j] = sum2
            pos_pair_accessible = (i, j)
            sum_close = fs.sums_close.get(pos_pair_accessible)
            if sum_close is not None:
                sum_accessible = fs.sums_accessible[pos_pair_accessible]
                sum_pair = (ZERO if i < 1 else fs.sums_external[0][i - 1],
                            ZERO if j > seq_len - 2 else fs.sums_external[j + 1][seq_len - 1])
                sum_ = sum_pair[0] + sum_accessible + sum_pair[1] - global_sum
                for k in reversed(range(0, i)):
                    if i - k - 1 > P.MAX_2LOOP_LEN:
                        break
                    for l in range(j + 1, seq_len):
                        if l - j - 1 + i - k - 1 > P.MAX_2LOOP_LEN:
                            break
                        pos_pair_close = (k, l)
                        x = fs.sums_close.get(pos_pair_close)
                        if x is not None:
                            sum_ = logsumexp(sum_, basepair_probs[pos_pair_close] + sum_close - x
                                             + fold_scores.twoloop_scores[(k, l, i, j)])
                sum_accessible = sum_accessible + P.COEFF_NUM_BRANCHES
                for k in range(0, i):
                    x = fs.sums_1ormore_basepairs[k + 1][i - 1]
                    sum_ = logsumexp(sum_, sum_accessible + probs_multibranch2[k][j] + x)
                    y = probs_multibranch[k][j]
                    sum_ = logsumexp(sum_, sum_accessible + y)
                    sum_ = logsumexp(sum_, sum_accessible + x + y)
                if sum_ > NEG_INFINITY:
                    basepair_probs[pos_pair_accessible] = sum_
    return {x: expf(y) for x, y in basepair_probs.items()}


def get_basepair_probs_contra(fs: FoldSums, seq_len: int, fold_scores: FoldScores, allows_short_hairpins: bool,
                              F: FoldScoreSets):
    """src/mccaskill_algo.rs:612-723"""
    global_sum = fs.sums_external[0][seq_len - 1]
    basepair_probs = {}
    probs_multibranch = [[NEG_INFINITY] * seq_len for _ in range(seq_len)]
    probs_multibranch2 = [[NEG_INFINITY] * seq_len for _ in range(seq_len)]
    for subseq_len in reversed(range(2 if allows_short_hairpins else F.MIN_SPAN_HAIRPIN_CLOSE, seq_len + 1)):
        for i in range(0, seq_len - subseq_len + 1):
            j = i + subseq_len - 1
            sum_ = NEG_INFINITY
            sum2 = sum_
            for k in range(j + 1, seq_len):
                pos_pair_close = (i, k)
                x = fs.sums_close.get(pos_pair_close)
                if x is not None:
                    basepair_prob = basepair_probs[pos_pair_close]
                    multibranch_close_score = fold_scores.multibranch_close_scores[pos_pair_close]
                    x = basepair_prob + multibranch_close_score - x
                    sum_ = logsumexp(sum_, x + fs.sums_1ormore_basepairs[j + 1][k - 1])
                    sum2 = logsumexp(sum2, x + F.multibranch_score_unpair * f32(k - j - 1))
            probs_multibranch[i][j] = sum_
            probs_multibranch2[i][j] = sum2
            pos_pair_accessible = (i, j)
            sum_close = fs.sums_close.get(pos_pair_accessible)
            if sum_close is not None:
                sum_pair = (ZERO if i < 1 else fs.sums_external[0][i - 1],
                            ZERO if j > seq_len - 2 else fs.sums_external[j + 1][seq_len - 1])
                sum_ = (sum_pair[0] + sum_pair[1] + fs.sums_accessible[pos_pair_accessible]
                        + F.external_score_basepair - global_sum)
                for k in reversed(range(0, i)):
                    if i - k - 1 > F.MAX_LOOP_LEN:
                        break
                    for l in range(j + 1, seq_len):
                        if l - j - 1 + i - k - 1 > F.MAX_LOOP_LEN:
                            break
                        pos_pair_close = (k, l)
                        x = fs.sums_close.get(pos_pair_close)
                        if x is not None:
                            sum_ = logsumexp(sum_, basepair_probs[pos_pair_close] + sum_close - x
                                             + fold_scores.twoloop_scores[(k, l, i, j)])
                sum_accessible = fs.sums_accessible[pos_pair_accessible] + F.multibranch_score_basepair
                for k in range(0, i):
                    x = fs.sums_1ormore_basepairs[k + 1][i - 1]
                    sum_ = logsumexp(sum_, sum_accessible + probs_multibranch2[k][j] + x)
                    y = probs_multibranch[k][j]
                    sum_ = logsumexp(sum_, sum_accessible + y + F.multibranch_score_unpair * f32(i - k - 1))
                    sum_ = logsumexp(sum_, sum_accessible + x + y)
                if sum_ > NEG_INFINITY:
                    basepair_probs[pos_pair_accessible] = sum_
    return {x: expf(y) for x, y in basepair_probs.items()}


def mccaskill_algo(seq, uses_contra_model: bool, allows_short_hairpins: bool, fold_score_sets, turner=None):
    """src/mccaskill_algo.rs:247-280 -> (SparseProbMat as a dict, FoldScores, FoldSums).
    `fold_score_sets`: RnaContraTables blob; `turner`: RnaTurnerTables blob (compile-time consts in the reference)."""
    seq = [int(b) for b in seq]
    seq_len = len(seq)
    fold_scores = FoldScores()
    if uses_contra_model:
        F = FoldScoreSets(fold_score_sets)
        fold_sums = get_fold_sums_contra(seq, fold_scores, allows_short_hairpins, F)
        basepair_probs = get_basepair_probs_contra(fold_sums, seq_len, fold_scores, allows_short_hairpins, F)
    else:
        P = Turner(turner)
        fold_sums = get_fold_sums(seq, fold_scores, P)
        basepair_probs = get_basepair_probs(fold_sums, seq_len, fold_scores, P)
    return basepair_probs, fold_scores, fold_sums


# ---------------------------------------------------------------------------------------------------------
# src/centroid_fold.rs:25-105
# ---------------------------------------------------------------------------------------------------------
def centroid_fold(basepair_probs: dict, seq_len: int, centroid_threshold):
    """-> (basepair_pos_pairs in traceback order, expect_accuracy)"""
    g = f32(centroid_threshold)
    ONE = f32(1.0)
    mea = [[ZERO] * seq_len for _ in range(seq_len)]
    for subseq_len in range(1, seq_len + 1):
        for i in range(0, seq_len - subseq_len + 1):
            j = i + subseq_len - 1
            if i == j:
                continue
            m = mea[i + 1][j]
            e = mea[i][j - 1]
            if e > m:
                m = e
            x = basepair_probs.get((i, j))
            if x is not None:
                e = mea[i + 1][j - 1] + g * x - ONE
                if e > m:
                    m = e
            for k in range(i + 1, j):
                e = mea[i][k] + mea[k + 1][j]
                if e > m:
                    m = e
            mea[i][j] = m
    pairs = []
    stack = [(0, seq_len - 1)]
    while stack:
        i, j = stack.pop()
        if j <= i:
            continue
        m = mea[i][j]
        if m == ZERO:
            continue
        if m == mea[i + 1][j]:
            stack.append((i + 1, j))
        elif m == mea[i][j - 1]:
            stack.append((i, j - 1))
        elif (i, j) in basepair_probs and m == mea[i + 1][j - 1] + g * basepair_probs[(i, j)] - ONE:
            stack.append((i + 1, j - 1))
            pairs.append((i, j))
        else:
            for k in range(i + 1, j):
                if m == mea[i][k] + mea[k + 1][j]:
                    stack.append((i, k))
                    stack.append((k + 1, j))
                    break
    return pairs, mea[0][seq_len - 1]


def get_fold_str(pairs, seq_len: int) -> str:
    s = ["."] * seq_len                              # src/bin/centroid_fold.rs:197-207
    for i, j in pairs:
        s[i] = "("
        s[j] = ")"
    return "".join(s)


# ---------------------------------------------------------------------------------------------------------
# src/durbin_algo.rs:59-242 (seq_pair is sentinel-padded: src/bin/durbin_algo.rs:48-50)
# ---------------------------------------------------------------------------------------------------------
class AlignScores:
    def __init__(self, a):
        for name in ("match2match_score", "match2insert_score", "insert_extend_score", "insert_switch_score",
                     "init_match_score", "init_insert_score"):
            setattr(self, name, f32(getattr(a, name)))
        self.insert_scores = _arr(a, "insert_scores")
        self.match_scores = _arr(a, "match_scores")


def durbin_algo(seq_pair, align_scores):
    s0 = [int(b) for b in seq_pair[0]]
    s1 = [int(b) for b in seq_pair[1]]
    S = AlignScores(align_scores)
    n, m = len(s0), len(s1)
    neg = lambda: [[NEG_INFINITY] * m for _ in range(n)]
    fm, fi, fd, bm, bi, bd = neg(), neg(), neg(), neg(), neg(), neg()
    for i in range(0, n - 1):
        for j in range(0, m - 1):
            if i == 0 and j == 0:
                fm[i][j] = ZERO
                continue
            if i > 0 and j > 0:
                sum_ = NEG_INFINITY
                match_score = S.match_scores[s0[i]][s1[j]]
                begins_sum = (i - 1, j - 1) == (0, 0)
                sum_ = logsumexp(sum_, fm[i - 1][j - 1] + (S.init_match_score if begins_sum else S.match2match_score))
                sum_ = logsumexp(sum_, fi[i - 1][j - 1] + S.match2insert_score)
                sum_ = logsumexp(sum_, fd[i - 1][j - 1] + S.match2insert_score)
                fm[i][j] = sum_ + match_score
            if i > 0:
                insert_score = S.insert_scores[s0[i]]
                begins_sum = (i - 1, j) == (0, 0)
                sum_ = NEG_INFINITY
                sum_ = logsumexp(sum_, fm[i - 1][j] + (S.init_insert_score if begins_sum else S.match2insert_score))
                sum_ = logsumexp(sum_, fi[i - 1][j] + S.insert_extend_score)
                fi[i][j] = sum_ + insert_score
            if j > 0:
                insert_score = S.insert_scores[s1[j]]
                begins_sum = (i, j - 1) == (0, 0)
                sum_ = NEG_INFINITY
                sum_ = logsumexp(sum_, fm[i][j - 1] + (S.init_insert_score if begins_sum else S.match2insert_score))
                sum_ = logsumexp(sum_, fd[i][j - 1] + S.insert_extend_score)
                fd[i][j] = sum_ + insert_score
    for i in reversed(range(1, n)):
        for j in reversed(range(1, m)):
            if i == n - 1 and j == m - 1:
                bm[i][j] = ZERO
                continue
            if i < n - 1 and j < m - 1:
                sum_ = NEG_INFINITY
                match_score = S.match_scores[s0[i]][s1[j]]
                ends_sum = (i + 1, j + 1) == (n - 1, m - 1)
                sum_ = logsumexp(sum_, bm[i + 1][j + 1] + (ZERO if ends_sum else S.match2match_score))
                sum_ = logsumexp(sum_, bi[i + 1][j + 1] + S.match2insert_score)
                sum_ = logsumexp(sum_, bd[i + 1][j + 1] + S.match2insert_score)
                bm[i][j] = sum_ + match_score
            if i < n - 1:
                insert_score = S.insert_scores[s0[i]]
                ends_sum = (i + 1, j) == (n - 1, m - 1)
                sum_ = NEG_INFINITY
                sum_ = logsumexp(sum_, bm[i + 1][j] + (ZERO if ends_sum else S.match2insert_score))
                sum_ = logsumexp(sum_, bi[i + 1][j] + S.insert_extend_score)
                bi[i][j] = sum_ + insert_score
            if j < m - 1:
                insert_score = S.insert_scores[s1[j]]
                ends_sum = (i, j + 1) == (n - 1, m - 1)
                sum_ = NEG_INFINITY
                sum_ = logsumexp(sum_, bm[i][j + 1] + (ZERO if ends_sum else S.match2insert_score))
                sum_ = logsumexp(sum_, bd[i][j + 1] + S.insert_extend_score)
                bd[i][j] = sum_ + insert_score
    # get_match_probs, src/durbin_algo.rs:201-242
    match_probs = [[ZERO] * m for _ in range(n)]
    global_sum = fm[n - 2][m - 2]
    global_sum = logsumexp(global_sum, fi[n - 2][m - 2])
    global_sum = logsumexp(global_sum, fd[n - 2][m - 2])
    for i in range(n):
        if i == 0 or i == n - 1:
            continue
        for j in range(m):
            if j == 0 or j == m - 1:
                continue
            sum_ = NEG_INFINITY
            forward_sum = fm[i][j]
            ends_sum = (i + 1, j + 1) == (n - 1, m - 1)
            sum_ = logsumexp(sum_, (ZERO if ends_sum else S.match2match_score) + bm[i + 1][j + 1])
            sum_ = logsumexp(sum_, S.match2insert_score + bi[i + 1][j + 1])
            sum_ = logsumexp(sum_, S.match2insert_score + bd[i + 1][j + 1])
            match_probs[i][j] = expf(forward_sum + sum_ - global_sum)
    return np.array(match_probs, dtype=np.float32)
