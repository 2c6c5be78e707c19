"""FAST numeric modes (include/rna_algos_b200.h RNA_NUMERIC_FAST_F32 / _F64; csrc/fast_kernel.cuh): the McCaskill
recurrences in exact log-space arithmetic, re-associated into warp-shuffle reductions.

Tolerances, against the EXACT-MATH oracle (oracle/oracle.c built with -DORC_EXACT: f64, log1p/exp):
  FAST_F64   logZ: 3e-7 relative; every BPP entry: 3e-7 absolute — the f32 rounding of the OUTPUTS (the north_star's
             "1e-5 relative in f64" with room)
  FAST_F32   logZ: 2e-6 relative; BPP: 2e-4 absolute (tRNA lengths) — the f32-path figure, stated separately
and against the reference's own approximate numerics (the bit-exact default mode): BPP within 2e-3 absolute — that
gap is the reference's polynomial logsumexp / expf error, not this kernel's."""
import numpy as np
import pytest

from common import default_tables, load_trnas, pack, random_seqs
from oracle_lib import Oracle
from rna_algos_b200 import tables as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def handle():
    from rna_algos_b200.api import Handle
    tt, ct, at = default_tables()
    h = Handle(0, tt, ct, at)
    yield h
    h.close()


@pytest.fixture(scope="module")
def exact():
    return Oracle(exact=True)


def run_mode(handle, mode, bases, offsets, contra, gammas, allows_short=False):
    handle.set_numeric_mode(mode)
    try:
        return handle.fold_batch(bases, offsets, contra, allows_short, gammas)
    finally:
        handle.set_numeric_mode("exact")


def errors(got, want):
    pg, pw = got["bpp"].astype(np.float64), want["bpp"].astype(np.float64)
    assert ((pg == T.BPP_ABSENT) == (pw == T.BPP_ABSENT)).all(), "different key sets"
    present = pw != T.BPP_ABSENT
    zw = want["logz"].astype(np.float64)
    dz = np.abs(got["logz"].astype(np.float64) - zw) / np.maximum(np.abs(zw), 1.0)   # (logZ = 0 for the shortest ones)
    return float(dz.max()), float(np.abs(pg[present] - pw[present]).max())


@pytest.mark.parametrize("contra", [False, True])
def test_fast_modes_on_trnas(handle, exact, contra):
    tt, ct, _ = default_tables()
    seqs = load_trnas() + random_seqs(3, [1, 2, 4, 5, 6, 17, 33, 64, 65, 120])
    bases, offsets = pack(seqs)
    gammas = [0.5, 2.0, 16.0]
    want = exact.fold_batch(bases, offsets, contra, False, tt, ct, gammas, n_threads=8)
    g64 = run_mode(handle, "fast64", bases, offsets, contra, gammas)
    ez, ep = errors(g64, want)
    assert ez < 3e-7 and ep < 3e-7, (ez, ep)   # (outputs are f32: 6e-8 relative per rounding)
    # the centroid estimator is the same exact max-plus code in every mode: structures = centroid_fold of these BPPs
    f32o = Oracle()
    for s_ in range(len(seqs)):
        L = len(seqs[s_])
        lo = int(g64["bpp_offsets"][s_])
        for gi, g in enumerate(gammas):
            st, _, ea = f32o.centroid(g64["bpp"][lo:lo + L * (L - 1) // 2], L, g)
            assert bytes(g64["structs"][gi, offsets[s_]:offsets[s_ + 1]]).decode() == st
            assert np.float32(ea) == g64["expect_acc"][gi, s_]
    g32 = run_mode(handle, "fast", bases, offsets, contra, gammas)
    ez, ep = errors(g32, want)
    assert ez < 2e-6 and ep < 2e-4, (ez, ep)
    # versus the reference's numerics (bit-exact mode): the reference's own approximation error
    ref = handle.fold_batch(bases, offsets, contra, False, gammas)
    ez, ep = errors(g32, ref)
    assert ep < 2e-3, ep


def test_fast_allows_short_hairpins_and_ragged(handle, exact):
    tt, ct, _ = default_tables()
    seqs = random_seqs(8, [3, 9, 30, 77, 150, 201])
    bases, offsets = pack(seqs)
    want = exact.fold_batch(bases, offsets, True, True, tt, ct, [1.0], n_threads=8)
    got = run_mode(handle, "fast64", bases, offsets, True, [1.0], allows_short=True)
    ez, ep = errors(got, want)
    assert ez < 3e-7 and ep < 3e-7, (ez, ep)   # (outputs are f32: 6e-8 relative per rounding)


@pytest.mark.parametrize("contra", [False, True])
def test_fast_long_sequence_cooperative(handle, exact, contra):
    """The cooperative (whole-GPU) FAST path: one 1500-nt sequence; f64 state against exact math, f32 stated."""
    import os
    tt, ct, _ = default_tables()
    seq = np.random.default_rng(12).integers(0, 4, size=1500).astype(np.uint8)
    bases, offsets = pack([seq])
    want = exact.fold_batch(bases, offsets, contra, False, tt, ct, [2.0], inner_threads=max(2, os.cpu_count() or 2))
    g64 = run_mode(handle, "fast64", bases, offsets, contra, [2.0])
    ez, ep = errors(g64, want)
    assert ez < 3e-7 and ep < 1e-6, (ez, ep)
    g32 = run_mode(handle, "fast", bases, offsets, contra, [2.0])
    ez, ep = errors(g32, want)
    assert ez < 5e-6 and ep < 5e-3, (ez, ep)


@pytest.mark.parametrize("contra", [False, True])
def test_fast_f32_mixed_batch_hybrid_dispatch(handle, exact, contra):
    """FAST_F32 sends the few long sequences of a call to fast_fold_kernel's cooperative grid (centroid from the packed
    BPPs, on that subset) and everything else to the FAST build of the batch kernel (shared-memory and HBM-resident
    modes, centroid in the kernel): one ragged call covers all of it.  State against exact math; structures must be
    centroid_fold of the returned BPPs whichever kernel produced them."""
    import os
    tt, ct, _ = default_tables()
    seqs = random_seqs(21, [1100, 400, 150, 89, 30, 5, 1, 260])
    bases, offsets = pack(seqs)
    gammas = [1.0, 8.0]
    want = exact.fold_batch(bases, offsets, contra, False, tt, ct, gammas, inner_threads=max(2, os.cpu_count() or 2))
    got = run_mode(handle, "fast", bases, offsets, contra, gammas)
    ez, ep = errors(got, want)
    assert ez < 5e-6 and ep < 5e-3, (ez, ep)
    f32o = Oracle()
    for s_ in range(len(seqs)):
        L = len(seqs[s_])
        lo = int(got["bpp_offsets"][s_])
        for gi, g in enumerate(gammas):
            st, _, ea = f32o.centroid(got["bpp"][lo:lo + L * (L - 1) // 2], L, g)
            assert bytes(got["structs"][gi, offsets[s_]:offsets[s_ + 1]]).decode() == st, (s_, L, g)
            assert np.float32(ea) == got["expect_acc"][gi, s_]
    # same call without the BPP output: the structures must not change (the cooperative part then stages its BPPs in
    # a buffer of the handle)
    handle.set_numeric_mode("fast")
    try:
        nob = handle.fold_batch(bases, offsets, contra, False, gammas, want_bpp=False)
    finally:
        handle.set_numeric_mode("exact")
    assert (nob["structs"] == got["structs"]).all() and (nob["logz"] == got["logz"]).all()
