"""Independent check of the oracle's recurrences: enumerate EVERY secondary structure of short
sequences, score each by its loop decomposition, and compare partition function and base-pairing
probabilities with the oracle's inside/outside (f64 exact-math flavour tightly; f32 reference-numerics
flavour within the reference's own approximation error).

This validates the restatement of src/mccaskill_algo.rs:282-723 (which structures are summed, and with
which loop scores) without access to a Rust toolchain.  It cannot validate upstream table VALUES.
"""
import ctypes as C
import math

import numpy as np
import pytest

from oracle_lib import Oracle, u8p
from rna_algos_b200 import tables as T


def _canon(x, y):
    return (x, y) in T.CANONICAL


class Model:
    def __init__(self, orc, seq, contra, allows_short, tt, ct):
        self.o, self.seq, self.contra, self.tt, self.ct = orc, np.ascontiguousarray(seq, dtype=np.uint8), contra, tt, ct
        self.L = len(seq)
        self.p = self.seq.ctypes.data_as(u8p)
        self.a = (C.byref(tt) if tt is not None else None, C.byref(ct) if ct is not None else None)
        if contra:
            self.minspan = 2 if allows_short else ct.min_span_hairpin_close
            self.max2 = ct.max_loop_len
        else:
            self.minspan = tt.min_span_hairpin_close
            self.max2 = tt.max_2loop_len

    def can_pair(self, i, j):
        return _canon(int(self.seq[i]), int(self.seq[j])) and (j - i + 1) >= self.minspan

    def hairpin(self, i, j):
        if self.contra and (j - i - 1) > self.ct.max_loop_len:
            return None
        return self.o.lib.orc_score_hairpin(self.p, self.L, i, j, int(self.contra), *self.a)

    def twoloop(self, i, j, k, l):
        if (k - i - 1) + (j - l - 1) > self.max2:
            return None
        return self.o.lib.orc_score_twoloop(self.p, self.L, i, j, k, l, int(self.contra), *self.a)

    def mbclose(self, i, j):
        return self.o.lib.orc_score_multibranch_close(self.p, self.L, i, j, int(self.contra), *self.a)

    def acc(self, i, j):
        return self.o.lib.orc_score_accessible(self.p, self.L, i, j, int(self.contra), *self.a)


def enumerate_structures(m: Model):
    """Yield (pairs, log_weight) for every structure the reference's grammar admits."""
    L = m.L

    def regions(i, j):
        """All sets of non-crossing pairs inside [i, j] (inclusive), as lists of top-level pairs with
        their recursively chosen interiors: yields (top_pairs, all_pairs)."""
        if j < i:
            yield [], []
            return
        # position i unpaired
        for tops, allp in regions(i + 1, j):
            yield tops, allp
        # position i paired with k
        for k in range(i + 1, j + 1):
            if not m.can_pair(i, k):
                continue
            for in_tops, in_all in regions(i + 1, k - 1):
                for r_tops, r_all in regions(k + 1, j):
                    yield [(i, k, in_tops)] + r_tops, [(i, k)] + in_all + r_all

    def pair_score(i, j, in_tops):
        """log-weight of everything closed by (i,j); None if not representable."""
        if len(in_tops) == 0:
            return m.hairpin(i, j)
        if len(in_tops) == 1:
            k, l, sub = in_tops[0]
            s = m.twoloop(i, j, k, l)
            if s is None:
                return None
            r = pair_score(k, l, sub)
            return None if r is None else s + r
        tot = m.mbclose(i, j)
        unp = (j - i - 1)
        for (k, l, sub) in in_tops:
            r = pair_score(k, l, sub)
            if r is None:
                return None
            unp -= (l - k + 1)
            if m.contra:
                tot += r + m.acc(k, l) + m.ct.multibranch_score_basepair
            else:
                tot += r + m.acc(k, l) + m.tt.coeff_num_branches
        if m.contra:
            tot += m.ct.multibranch_score_unpair * unp
        return tot

    for tops, allp in regions(0, L - 1):
        tot = 0.0
        unp = L
        ok = True
        for (k, l, sub) in tops:
            r = pair_score(k, l, sub)
            if r is None:
                ok = False
                break
            unp -= (l - k + 1)
            tot += r + m.acc(k, l)
            if m.contra:
                tot += m.ct.external_score_basepair
        if not ok:
            continue
        if m.contra:
            tot += m.ct.external_score_unpair * unp
        yield allp, tot


def brute(m: Model):
    L = m.L
    items = list(enumerate_structures(m))
    mx = max(w for _, w in items)
    z = sum(math.exp(w - mx) for _, w in items)
    logz = mx + math.log(z)
    bpp = np.zeros((L, L))
    for pairs, w in items:
        pw = math.exp(w - logz)
        for (i, j) in pairs:
            bpp[i, j] += pw
    return logz, bpp, len(items)


def _biased_seq(rng, L):
    # G/C-rich hairpin-prone sequences so that multiloops and two-loops actually occur
    return rng.choice(4, size=L, p=[0.2, 0.3, 0.3, 0.2]).astype(np.uint8)


CASES = [
    # (seed, L, contra, allows_short)
    (1, 11, False, False), (2, 13, False, False), (3, 15, False, False), (4, 16, False, False),
    (5, 11, True, False), (6, 13, True, False), (7, 15, True, False), (8, 16, True, False),
    (9, 10, True, True), (10, 12, True, True), (15, 14, True, True),
    (11, 20, False, False), (12, 20, True, False), (13, 22, False, False), (14, 22, True, False),
]


@pytest.mark.parametrize("seed,L,contra,allows_short", CASES)
def test_recurrences_match_enumeration(seed, L, contra, allows_short):
    rng = np.random.default_rng(1000 + seed)
    seq = _biased_seq(rng, L)
    if contra:
        tt, ct = None, T.random_contra_tables(seed)
        ct.min_span_hairpin_close = 4
    else:
        tt, ct = T.random_turner_tables(seed), None
        tt.min_span_hairpin_close = 4
        tt.min_hairpin_len = 2
        tt.max_2loop_len = int(rng.integers(2, 6))
    ox = Oracle(exact=True)
    o32 = Oracle(exact=False)
    m = Model(ox, seq, contra, allows_short, tt, ct)
    logz, bpp, n_structs = brute(m)
    assert n_structs > 1
    got, got_logz = ox.mccaskill(seq, contra, allows_short, tt, ct)
    assert abs(float(got_logz) - logz) <= 1e-5 * max(1.0, abs(logz))  # logz returned as f32
    got32, got32_logz = o32.mccaskill(seq, contra, allows_short, tt, ct)
    assert abs(float(got32_logz) - logz) <= 2e-3
    for i in range(L):
        for j in range(i + 1, L):
            v = got[i * (2 * L - i - 1) // 2 + (j - i - 1)]
            v32 = got32[i * (2 * L - i - 1) // 2 + (j - i - 1)]
            if bpp[i, j] == 0.0:
                # never formed in any admissible structure
                assert v == T.BPP_ABSENT or abs(v) < 1e-6
            else:
                assert v != T.BPP_ABSENT
                assert abs(v - bpp[i, j]) <= 2e-6 + 1e-5 * bpp[i, j], (i, j, v, bpp[i, j])
                # f32 + CONTRAfold polynomials: the reference's own approximation error
                assert abs(v32 - bpp[i, j]) <= 5e-3, (i, j, v32, bpp[i, j])


def test_real_tables_small():
    """Same check with the restated Turner / CONTRAfold blobs (real caps: span >= 5, loops <= 30)."""
    rng = np.random.default_rng(77)
    ox = Oracle(exact=True)
    for contra in (False, True):
        seq = _biased_seq(rng, 17)
        tt, ct = T.standin_turner_tables(), T.standin_contra_tables()
        m = Model(ox, seq, contra, False, tt, ct)
        logz, bpp, n = brute(m)
        got, got_logz = ox.mccaskill(seq, contra, False, tt, ct)
        assert abs(float(got_logz) - logz) <= 1e-5 * max(1.0, abs(logz))
        L = len(seq)
        for i in range(L):
            for j in range(i + 1, L):
                v = got[i * (2 * L - i - 1) // 2 + (j - i - 1)]
                if bpp[i, j] > 0:
                    assert abs(v - bpp[i, j]) <= 2e-6 + 1e-5 * bpp[i, j]
