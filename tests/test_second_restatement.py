"""Two independent restatements of the reference must agree bit for bit.

oracle/oracle.c (C, dense arrays) and oracle/translit.py (Python, a literal transliteration of the Rust text with
the reference's own hash-map data structures) were written separately from /root/reference/src/*.rs.  The reference
cannot be built here and pins no numbers itself (tests/tests.rs:33,38,74), so agreement of two readings — on the
bundled tRNAs (tests/tests.rs:7-43's inputs), both models, the centroid estimator and the pair-HMM, plus random
sequences with random score tables — is the strongest evidence available that the oracle follows the reference.
CPU only."""
import numpy as np
import pytest

import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import translit as TL  # noqa: E402

from common import assert_bits_equal, default_tables, load_trnas, random_seqs
from oracle_lib import Oracle
from rna_algos_b200 import tables as T


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def packed(bp: dict, L: int) -> np.ndarray:
    out = np.full(L * (L - 1) // 2, T.BPP_ABSENT, dtype=np.float32)
    for (i, j), p in bp.items():
        out[i * (2 * L - i - 1) // 2 + (j - i - 1)] = p
    return out


def check_fold(oracle, seq, contra, allows_short, tt, ct, gammas):
    L = len(seq)
    bp, scores, sums = TL.mccaskill_algo(seq, contra, allows_short, ct, tt)
    want_bpp, want_logz, dbg = oracle.mccaskill(seq, contra, allows_short, tt, ct, debug=True)
    assert_bits_equal(packed(bp, L), want_bpp, f"BPP L={L} contra={contra}")
    assert_bits_equal(np.float32(sums.sums_external[0][L - 1]), np.float32(want_logz), "logZ")
    # the reference's own (only) assertion: tests/tests.rs:33,38
    assert all(-0.001 <= float(p) < 1.001 for p in bp.values())
    # inside state: sums_close keys / values, sums_external, sums_1ormore_basepairs
    close = np.full((L, L), -np.inf, dtype=np.float32)
    for (i, j), v in sums.sums_close.items():
        close[i, j] = v
    assert_bits_equal(close, dbg["close"], "sums_close")
    assert_bits_equal(np.array(sums.sums_external, dtype=np.float32), dbg["external"], "sums_external")
    assert_bits_equal(np.array(sums.sums_1ormore_basepairs, dtype=np.float32), dbg["m1"], "sums_1ormore_basepairs")
    for g in gammas:
        pairs, ea = TL.centroid_fold(bp, L, g)
        s, wp, wea = oracle.centroid(want_bpp, L, g)
        assert TL.get_fold_str(pairs, L) == s, f"structure gamma={g}"
        assert [tuple(int(v) for v in p) for p in wp] == pairs, "traceback order"
        assert_bits_equal(np.float32(ea), np.float32(wea), "expect_accuracy")
    return scores


@pytest.mark.parametrize("contra", [False, True])
def test_trnas_both_restatements_agree(oracle, contra):
    tt, ct, _ = default_tables()
    for seq in load_trnas():
        check_fold(oracle, seq, contra, False, tt, ct, [2.0 ** -3, 1.0, 2.0, 64.0])


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_tables_and_sequences(oracle, seed):
    rng = np.random.default_rng(900 + seed)
    tt = T.random_turner_tables(seed, special=True)
    ct = T.random_contra_tables(seed)
    lens = [int(x) for x in rng.integers(1, 46, size=5)] + [1, 2, 5]
    for seq in random_seqs(seed, lens):
        for contra in (False, True):
            check_fold(oracle, seq, contra, bool(contra and seed % 2), tt, ct, [0.5, 4.0])


def test_fold_scores_memo_matches_the_oracle_scorers(oracle):
    """FoldScores<T> (src/mccaskill_algo.rs:13-22): the per-pair memo maps of the transliteration against the
    oracle's scorer functions, key by key (both models)."""
    import ctypes as C
    tt, ct, _ = default_tables()
    seq = load_trnas()[4]
    L = len(seq)
    sp = np.ascontiguousarray(seq, dtype=np.uint8).ctypes.data_as(C.POINTER(C.c_uint8))
    for contra in (False, True):
        scores = TL.mccaskill_algo(seq, contra, False, ct, tt)[1]
        lib = oracle.lib
        for (i, j), v in scores.hairpin_scores.items():
            assert np.float32(lib.orc_score_hairpin(sp, L, i, j, int(contra), C.byref(tt), C.byref(ct))) == v
        for (i, j), v in scores.multibranch_close_scores.items():
            assert np.float32(lib.orc_score_multibranch_close(sp, L, i, j, int(contra), C.byref(tt), C.byref(ct))) == v
        for (i, j), v in scores.accessible_scores.items():
            assert np.float32(lib.orc_score_accessible(sp, L, i, j, int(contra), C.byref(tt), C.byref(ct))) == v
        for n, ((i, j, k, l), v) in enumerate(scores.twoloop_scores.items()):
            if n % 7 == 0:
                assert np.float32(lib.orc_score_twoloop(sp, L, i, j, k, l, int(contra), C.byref(tt), C.byref(ct))) == v
        assert len(scores.twoloop_scores) > 10000


def test_durbin_both_restatements_agree(oracle):
    _, _, at = default_tables()
    seqs = load_trnas()
    pad = lambda s: [TL.PSEUDO_BASE] + [int(b) for b in s] + [TL.PSEUDO_BASE]   # src/bin/durbin_algo.rs:48-50
    for a, b in ((4, 2), (0, 5)):
        got = TL.durbin_algo((pad(seqs[a]), pad(seqs[b])), at)
        want = oracle.durbin(seqs[a], seqs[b], at)
        assert_bits_equal(got, want, f"match probabilities pair {a},{b}")
        inner = got[1:-1, 1:-1]
        assert (inner >= -0.001).all() and (inner < 1.001).all()       # tests/tests.rs:74
        assert (got[0] == 0).all() and (got[:, 0] == 0).all()
    rs = random_seqs(77, [1, 2, 9])
    rat = T.random_align_tables(5)
    for a in rs:
        for b in rs:
            assert_bits_equal(TL.durbin_algo((pad(a), pad(b)), rat), oracle.durbin(a, b, rat), "random pair")


@pytest.mark.parametrize("contra", [False, True])
def test_fold_sums_planes_of_both_restatements(oracle, contra):
    """FoldSums + FoldScores member by member (the oracle's orc_fold_sums planes are what the CUDA path's
    rna_fold_sums_batch is compared with in tests/test_gpu_parity.py)."""
    tt, ct, _ = default_tables()
    seq = load_trnas()[4]
    L = len(seq)
    _, scores, sums = TL.mccaskill_algo(seq, contra, False, ct, tt)
    planes = oracle.fold_sums(seq, contra, False, tt, ct)

    def sparse(d):
        m = np.full((L, L), -np.inf, dtype=np.float32)
        for (i, j), v in d.items():
            m[i, j] = v
        return m
    iu = np.triu_indices(L)
    want = [sparse(sums.sums_close), sparse(sums.sums_accessible), np.array(sums.sums_external, dtype=np.float32),
            np.array(sums.sums_rightmost_basepairs_external, dtype=np.float32),
            np.array(sums.sums_rightmost_basepairs_multibranch, dtype=np.float32),
            np.array(sums.sums_multibranch, dtype=np.float32), np.array(sums.sums_1ormore_basepairs, dtype=np.float32),
            sparse(scores.hairpin_scores), sparse(scores.multibranch_close_scores), sparse(scores.accessible_scores)]
    for p, w in enumerate(want):
        assert_bits_equal(planes[p][iu], w[iu], f"plane {p}")


@pytest.mark.parametrize("contra", [False, True])
def test_twoloop_scores_memo_of_both_restatements(oracle, contra):
    """FoldScores::twoloop_scores (the 4-D memo): the oracle-side enumeration that the CUDA export is compared with
    (tests/oracle_lib.py twoloop_scores) against the transliteration's dict — same keys, same f32 values."""
    tt, ct, _ = default_tables()
    for seq, short in ((load_trnas()[3], False), (random_seqs(3, [40])[0], True)):
        scores = TL.mccaskill_algo(seq, contra, short, ct, tt)[1]
        lst = oracle.twoloop_scores(seq, contra, short, tt, ct)
        assert len(lst) == len(scores.twoloop_scores) and len(lst) > 100
        for (i, j, k, l, sc) in lst:
            assert np.float32(scores.twoloop_scores[(i, j, k, l)]).view(np.uint32) == sc.view(np.uint32), (i, j, k, l)
