"""Parity of the CUDA path (through the C ABI) against the CPU oracle: BIT-EXACT for partition
functions, base-pairing probabilities, expected accuracies, dot-bracket structures and match
probabilities (all f32; tolerance 0 ULP)."""
import numpy as np
import pytest

from common import assert_bits_equal, default_tables, load_trnas, pack, random_seqs
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
def oracle():
    return Oracle()


SWEEP = [float(np.float32(2.0) ** p) for p in range(-7, 11)]


def check_fold(handle, oracle, seqs, contra, allows_short, gammas, tt, ct, threads=8):
    bases, offsets = pack(seqs)
    got = handle.fold_batch(bases, offsets, contra, allows_short, gammas)
    want = oracle.fold_batch(bases, offsets, contra, allows_short, tt, ct, gammas, n_threads=threads)
    assert_bits_equal(got["logz"], want["logz"], "logZ")
    assert_bits_equal(got["bpp"], want["bpp"], "BPP")
    assert_bits_equal(got["expect_acc"], want["expect_acc"], "expect_accuracy")
    assert (got["structs"] == want["structs"]).all(), "dot-bracket structures differ"
    return got, want


@pytest.mark.parametrize("contra", [False, True])
def test_trnas_bit_exact(handle, oracle, contra):
    """Config 1/2 parity set: the 6 bundled tRNAs, Turner and CONTRAfold, gamma = 1 and the sweep."""
    tt, ct, _ = default_tables()
    got, _ = check_fold(handle, oracle, load_trnas(), contra, False, [1.0] + SWEEP, tt, ct)
    p = got["bpp"]
    present = p != T.BPP_ABSENT
    # the reference's own (only) assertion: tests/tests.rs:33,38
    assert (p[present] >= -0.001).all() and (p[present] < 1.001).all()
    assert (got["structs"] == ord("(")).sum() > 0


@pytest.mark.parametrize("contra", [False, True])
def test_edge_lengths(handle, oracle, contra):
    """L = 1 .. 40: below / at / above MIN_SPAN_HAIRPIN_CLOSE, empty structures, ragged batch."""
    tt, ct, _ = default_tables()
    seqs = random_seqs(11, list(range(1, 41)))
    check_fold(handle, oracle, seqs, contra, False, [0.5, 2.0, 1024.0], tt, ct)


def test_allows_short_hairpins(handle, oracle):
    tt, ct, _ = default_tables()
    seqs = random_seqs(12, [2, 3, 4, 5, 9, 17, 33, 60])
    check_fold(handle, oracle, seqs, True, True, [2.0, 16.0], tt, ct)


@pytest.mark.parametrize("contra", [False, True])
def test_random_batch_ragged(handle, oracle, contra):
    """Ragged batch crossing several shared-memory buckets, incl. the 2-loop cap (spans > 32)."""
    tt, ct, _ = default_tables()
    rng = np.random.default_rng(5)
    seqs = random_seqs(13, rng.integers(20, 140, size=48))
    check_fold(handle, oracle, seqs, contra, False, [1.0, 4.0], tt, ct)


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("contra", [False, True])
def test_random_tables_fuzz(oracle, seed, contra):
    """Arbitrary tables (caps included): parity does not depend on the restated table values."""
    from rna_algos_b200.api import Handle
    tt = T.random_turner_tables(100 + seed)
    ct = T.random_contra_tables(200 + seed)
    h = Handle(0, tt, ct, None)
    try:
        seqs = random_seqs(20 + seed, [5, 8, 13, 21, 34, 55, 76, 89, 100])
        # hairpin-prone sequences exercise the special-hairpin list and the interior tables
        rng = np.random.default_rng(seed)
        seqs += [rng.choice(4, size=64, p=[0.15, 0.35, 0.35, 0.15]).astype(np.uint8) for _ in range(4)]
        check_fold(h, oracle, seqs, contra, False, [1.0, 8.0], tt, ct)
    finally:
        h.close()


def test_mid_length_global_path(handle, oracle):
    """Rfam-length sequences: matrices in the HBM/L2 workspace (one CTA per sequence)."""
    tt, ct, _ = default_tables()
    seqs = random_seqs(31, [150, 201, 333, 420])
    for contra in (False, True):
        check_fold(handle, oracle, seqs, contra, False, [2.0], tt, ct)


def test_long_cooperative_path(handle, oracle):
    """> 1024 nt: multi-CTA per-diagonal wavefront with a grid-wide barrier."""
    tt, ct, _ = default_tables()
    seqs = random_seqs(32, [1100])
    check_fold(handle, oracle, seqs, False, False, [2.0], tt, ct)


def test_single_item_api(handle, oracle):
    """The reference's per-call granularity: mccaskill_algo / centroid_fold on one sequence."""
    tt, ct, _ = default_tables()
    seq = load_trnas()[4]
    for contra in (False, True):
        bpp, logz = handle.mccaskill_algo(seq, contra, False)
        wbpp, wlogz = oracle.mccaskill(seq, contra, False, tt, ct)
        assert_bits_equal(bpp, wbpp, "BPP")
        assert_bits_equal(np.array([logz]), np.array([wlogz]), "logZ")
        for g in (0.25, 1.0, 2.0, 64.0):
            s, pairs, ea = handle.centroid_fold(bpp, len(seq), g)
            ws, wpairs, wea = oracle.centroid(wbpp, len(seq), g)
            assert s == ws
            assert (pairs == wpairs).all(), "traceback order of basepair_pos_pairs differs"
            assert_bits_equal(np.array([ea]), np.array([wea]), "expect_accuracy")


def test_centroid_batch_from_host_bpp(handle, oracle):
    tt, ct, _ = default_tables()
    seqs = load_trnas() + random_seqs(40, [30, 200])
    bases, offsets = pack(seqs)
    want = oracle.fold_batch(bases, offsets, False, False, tt, ct, SWEEP, n_threads=8)
    got = handle.centroid_batch(want["bpp"], offsets, SWEEP)
    assert (got["structs"] == want["structs"]).all()
    assert_bits_equal(got["expect_acc"], want["expect_acc"], "expect_accuracy")


def test_durbin_trna_pairs(handle, oracle):
    """Config 5 parity set: all 15 tRNA pairs with the genuine CONTRAlign v2.01 scores."""
    _, _, at = default_tables()
    seqs = load_trnas()
    bases, offsets = pack(seqs)
    pairs = np.array([(a, b) for a in range(6) for b in range(a + 1, 6)], dtype=np.uint32)
    got = handle.durbin_batch(bases, offsets, pairs)
    want = oracle.durbin_batch(bases, offsets, pairs, at, n_threads=8)
    assert_bits_equal(got["probs"], want["probs"], "match probs")
    p = got["probs"]
    assert (p >= -0.001).all() and (p < 1.001).all()   # tests/tests.rs:74


def test_durbin_random(handle, oracle):
    rng = np.random.default_rng(9)
    seqs = random_seqs(50, [1, 2, 3, 7, 40, 130, 300, 77])
    bases, offsets = pack(seqs)
    pairs = np.array([(a, b) for a in range(8) for b in range(8) if a != b], dtype=np.uint32)
    at = T.random_align_tables(3)
    from rna_algos_b200.api import Handle
    h = Handle(0, None, None, at)
    try:
        got = h.durbin_batch(bases, offsets, pairs)
    finally:
        h.close()
    want = oracle.durbin_batch(bases, offsets, pairs, at, n_threads=8)
    assert_bits_equal(got["probs"], want["probs"], "match probs")


def test_errors_do_not_abort(handle):
    from rna_algos_b200.api import RnaError
    with pytest.raises(RnaError):
        handle.fold_batch(np.array([0, 1, 7], dtype=np.uint8), np.array([0, 3], dtype=np.uint32), False)
    with pytest.raises(RnaError):
        handle.fold_batch(np.zeros(0, dtype=np.uint8), np.array([0, 0], dtype=np.uint32), False)


@pytest.mark.parametrize("contra", [False, True])
def test_golden_vectors(handle, contra):
    """The committed golden vectors (tests/golden/trna_oracle.npz, made by tests/golden/make_golden.py)."""
    import os
    from common import ROOT
    g = np.load(os.path.join(ROOT, "tests", "golden", "trna_oracle.npz"))
    bases, offsets = pack(load_trnas())
    got = handle.fold_batch(bases, offsets, contra, False, g["gammas"].tolist())
    k = "contra" if contra else "turner"
    assert_bits_equal(got["logz"], g[k + "_logz"], "logZ")
    assert_bits_equal(got["bpp"], g[k + "_bpp"], "BPP")
    assert_bits_equal(got["expect_acc"], g[k + "_expect_acc"], "expect_accuracy")
    assert (got["structs"] == g[k + "_structs"]).all()


@pytest.mark.parametrize("contra", [False, True])
def test_full_bench_size_properties(handle, oracle, contra):
    """BASELINE configs[0]/[1] at bench size (24 576 tiled tRNAs): every copy of a sequence must reproduce the
    oracle's result for that sequence bit for bit (determinism across CTAs, work-queue order and stream slots)."""
    tt, ct, _ = default_tables()
    base = load_trnas()
    n = 24576
    seqs = [base[i % 6] for i in range(n)]
    bases, offsets = pack(seqs)
    got = handle.fold_batch(bases, offsets, contra, False, [1.0, 2.0])
    b6, o6 = pack(base)
    want = oracle.fold_batch(b6, o6, contra, False, tt, ct, [1.0, 2.0], n_threads=6)
    lz = got["logz"].view(np.uint32).reshape(-1, 6)
    assert (lz == want["logz"].view(np.uint32)[None, :]).all(), "logZ differs between copies"
    ea = got["expect_acc"].view(np.uint32).reshape(2, -1, 6)
    assert (ea == want["expect_acc"].view(np.uint32)[:, None, :]).all()
    bo, wo = got["bpp_offsets"].astype(np.int64), want["bpp_offsets"].astype(np.int64)
    for s in range(6):
        ref = want["bpp"][wo[s]:wo[s + 1]].view(np.uint32)
        L = len(base[s])
        idx = np.arange(s, n, 6)
        blk = np.stack([got["bpp"][bo[k]:bo[k + 1]] for k in idx[:: max(1, len(idx) // 64)]]).view(np.uint32)
        assert (blk == ref[None, :]).all(), f"BPP of tRNA {s} differs between copies"
        st = got["structs"][:, offsets[idx[-1]]:offsets[idx[-1]] + L]
        assert (st == want["structs"][:, o6[s]:o6[s] + L]).all()
    p = got["bpp"][got["bpp"] != T.BPP_ABSENT]
    assert (p >= -0.001).all() and (p < 1.001).all()            # the reference's range assertion at full size


def test_rfam_like_lengths(handle, oracle):
    """BASELINE configs[2] stand-in: ragged 50..500-nt batch across shared-memory and HBM-resident buckets."""
    tt, ct, _ = default_tables()
    rng = np.random.default_rng(77)
    lens = np.exp(rng.uniform(np.log(50), np.log(500), size=24)).astype(int)
    check_fold(handle, oracle, random_seqs(78, lens), True, False, [1.0, 4.0], tt, ct)
    check_fold(handle, oracle, random_seqs(79, lens[:10]), False, False, [2.0], tt, ct)


def test_skewed_composition_falls_back(handle, oracle):
    """GU-only sequences make ~half of all cells closable: their term streams overflow the slot and the kernel
    must fall back to scoring on the fly, with identical results."""
    tt, ct, _ = default_tables()
    rng = np.random.default_rng(3)
    seqs = [rng.choice(np.array([2, 3], dtype=np.uint8), size=L) for L in (60, 90, 120)]
    check_fold(handle, oracle, seqs, True, False, [2.0], tt, ct)
    check_fold(handle, oracle, seqs, False, False, [2.0], tt, ct)


def test_long_cooperative_contra(handle, oracle):
    """> 1024 nt under CONTRAfold: grid-wide one-diagonal wavefront with roles spread over the SMs."""
    tt, ct, _ = default_tables()
    check_fold(handle, oracle, random_seqs(33, [1300]), True, False, [2.0, 0.5], tt, ct)


@pytest.mark.parametrize("L,seed", [(1024, 1), (2048, 2), (4096, 3)])
@pytest.mark.parametrize("contra", [False, True])
def test_config3_long_single_sequences(handle, oracle, L, seed, contra):
    """BASELINE configs[3]: one i.i.d.-uniform sequence at 1024 / 2048 / 4096 nt, seeds 1/2/3,
    both models, through the cooperative multi-CTA wavefront: logZ, every BPP entry, expected accuracies and
    structures bit-equal to the oracle (whose span loops run on all host cores here: same folds, same order)."""
    import os
    tt, ct, _ = default_tables()
    seq = np.random.default_rng(seed).integers(0, 4, size=L).astype(np.uint8)
    bases, offsets = pack([seq])
    gammas = [1.0, 2.0, 8.0]
    got = handle.fold_batch(bases, offsets, contra, False, gammas)
    want = oracle.fold_batch(bases, offsets, contra, False, tt, ct, gammas, inner_threads=max(2, os.cpu_count() or 2))
    assert_bits_equal(got["logz"], want["logz"], f"logZ L={L}")
    assert_bits_equal(got["bpp"], want["bpp"], f"BPP L={L}")
    assert_bits_equal(got["expect_acc"], want["expect_acc"], f"expect_accuracy L={L}")
    assert (got["structs"] == want["structs"]).all(), f"dot-bracket structures differ at L={L}"
    assert (got["structs"][1] == ord("(")).sum() > 10   # gamma = 2: a real structure, not all dots


def test_mixed_long_and_mid_routing(handle, oracle):
    """A few long sequences next to mid-length ones: the cost model sends the longest to the cooperative kernel and
    the rest to the one-CTA HBM-resident mode; results are bit-identical either way."""
    tt, ct, _ = default_tables()
    seqs = random_seqs(91, [1100, 700, 640, 300, 280, 260, 120, 76, 30])
    check_fold(handle, oracle, seqs, True, False, [1.0, 2.0], tt, ct)
    check_fold(handle, oracle, seqs[1:], False, False, [2.0], tt, ct)


@pytest.mark.parametrize("contra", [False, True])
def test_several_long_sequences_run_on_concurrent_cooperative_grids(handle, oracle, contra):
    """Several > 1024-nt sequences in one call: the launcher runs up to four cooperative grids side by side, each on its
    share of the SMs with scratch of its own.  Bit parity with the oracle on two of them, and every sequence
    bit-identical to the same sequence folded alone (the whole-GPU grid that test_config3 pins against the oracle)."""
    import os
    tt, ct, _ = default_tables()
    seqs = random_seqs(64, [1100, 1090, 1061, 1040, 1030, 1026, 700, 90])
    bases, offsets = pack(seqs)
    gammas = [1.0, 4.0]
    got = handle.fold_batch(bases, offsets, contra, False, gammas)
    bo = got["bpp_offsets"]
    for s_ in range(len(seqs)):
        b1, o1 = pack([seqs[s_]])
        if s_ in (1, 4):
            alone = oracle.fold_batch(b1, o1, contra, False, tt, ct, gammas, inner_threads=max(2, os.cpu_count() or 2))
        else:
            alone = handle.fold_batch(b1, o1, contra, False, gammas)
        assert_bits_equal(got["logz"][s_:s_ + 1], alone["logz"], f"logZ seq {s_}")
        assert_bits_equal(got["bpp"][int(bo[s_]):int(bo[s_ + 1])], alone["bpp"], f"BPP seq {s_}")
        assert_bits_equal(got["expect_acc"][:, s_], alone["expect_acc"][:, 0], f"expect_accuracy seq {s_}")
        assert (got["structs"][:, offsets[s_]:offsets[s_ + 1]] == alone["structs"]).all(), f"structures seq {s_}"


@pytest.mark.parametrize("seed,contra", [(301, True), (302, False), (303, True)])
def test_random_ragged_batches_through_every_route(handle, oracle, seed, contra):
    """Seeded ragged batches whose lengths straddle every routing boundary of the launcher at once — the 4-nt
    shared-memory buckets, the 96-nt classes of the HBM-resident mode, the cooperative / one-CTA cost model and the
    concurrent cooperative grids — in shuffled order: results must be bit-identical whatever route a sequence takes."""
    import os
    tt, ct, _ = default_tables()
    rng = np.random.default_rng(seed)
    lens = [int(x) for x in rng.integers(1, 230, size=24)] + [int(x) for x in rng.integers(221, 720, size=10 if seed != 303 else 3)]
    if seed == 303:
        lens += [1040, 1031]
    rng.shuffle(lens)
    seqs = random_seqs(seed, lens)
    check_fold(handle, oracle, seqs, contra, bool(contra and seed == 303), [1.0, 6.0], tt, ct, threads=max(8, os.cpu_count() or 8))


def test_long_cooperative_edge_lengths(handle, oracle):
    """The cooperative kernel is also the route for a lone mid-length sequence: odd/even lengths around the warp and
    pair-step boundaries (the last pair step handles one or two diagonals)."""
    tt, ct, _ = default_tables()
    for L in (384, 385, 415, 416, 449):
        for contra in (False, True):
            check_fold(handle, oracle, random_seqs(100 + L, [L]), contra, False, [2.0], tt, ct)


def test_centroid_batch_long_sequence(handle, oracle):
    """rna_centroid_batch on a > 1024-nt BPP matrix: the whole-grid chunked max-plus fill (atomic integer max of
    non-negative floats) must reproduce the reference's structures and expected accuracies."""
    L = 1200
    rng = np.random.default_rng(17)
    bpp = np.full(L * (L - 1) // 2, -1.0, dtype=np.float32)
    idx = rng.choice(bpp.shape[0], size=bpp.shape[0] // 5, replace=False)
    bpp[idx] = (rng.uniform(0, 1, size=idx.shape[0]) ** 3).astype(np.float32)
    offsets = np.array([0, L], dtype=np.uint32)
    gammas = [1.0, 4.0, 64.0]
    got = handle.centroid_batch(bpp, offsets, gammas)
    for g, gamma in enumerate(gammas):
        s, _, ea = oracle.centroid(bpp, L, gamma)
        assert bytes(got["structs"][g]).decode() == s
        assert np.float32(got["expect_acc"][g, 0]).view(np.uint32) == np.float32(ea).view(np.uint32)


def test_too_long_for_32bit_offsets_is_an_error(handle):
    """Beyond 46340 nt the cooperative kernel's 32-bit matrix offsets would overflow: the call must fail with
    RNA_ERR_TOO_LONG before touching the device, not corrupt memory."""
    from rna_algos_b200.api import RnaError
    seq = np.zeros(46341, dtype=np.uint8)
    with pytest.raises(RnaError) as e:
        handle.fold_batch(seq, np.array([0, 46341], dtype=np.uint32), False, False, [1.0])
    assert e.value.code == 4 and "46340" in str(e.value)   # RNA_ERR_TOO_LONG


def test_cooperative_random_lengths_stress(handle, oracle):
    """The cooperative kernel synchronises warps four ways (grid barrier, named producer/consumer barrier, CTA barriers,
    atomic-max centroid): a spread of lengths — odd and even, around warp (32) and batch (32 split points) multiples —
    one sequence per call so that each takes the cooperative route, each run twice to catch order-dependent results."""
    tt, ct, _ = default_tables()
    rng = np.random.default_rng(4242)
    lens = [384, 415, 448, 449, 511, 512, 513, 640, 777, 1025]
    for L in lens:
        contra = bool(rng.integers(0, 2))
        seq = rng.integers(0, 4, size=L).astype(np.uint8)
        bases, offsets = pack([seq])
        want = oracle.fold_batch(bases, offsets, contra, False, tt, ct, [0.5, 4.0], n_threads=8)
        for _ in range(2):
            got = handle.fold_batch(bases, offsets, contra, False, [0.5, 4.0])
            assert_bits_equal(got["bpp"], want["bpp"], f"BPP L={L} contra={contra}")
            assert_bits_equal(got["logz"], want["logz"], f"logZ L={L}")
            assert (got["structs"] == want["structs"]).all(), L
            assert_bits_equal(got["expect_acc"], want["expect_acc"], f"expected accuracy L={L}")


def test_zero_copy_bpp_into_pinned_host_buffer(handle, oracle):
    """A page-locked output buffer is written by the kernels directly (no staging copy); a pageable one is staged.
    Both must hold the same bits, in the batch path and in the cooperative path, and a buffer that is only partly
    page-locked must take the staged path."""
    import torch
    tt, ct, _ = default_tables()
    for seqs in (load_trnas() + random_seqs(55, [40, 120, 300]), random_seqs(56, [600])):
        bases, offsets = pack(seqs)
        want = handle.fold_batch(bases, offsets, True, False, [2.0])                 # pageable numpy outputs
        nbpp = want["bpp"].shape[0]
        pinned = torch.full((nbpp + 8,), 7.0, dtype=torch.float32).pin_memory().numpy()
        got = handle.fold_batch(bases, offsets, True, False, [2.0], out={"bpp": pinned[:nbpp]})
        assert got["bpp"] is not None and np.shares_memory(got["bpp"], pinned)
        assert_bits_equal(got["bpp"], want["bpp"], "zero-copy BPP")
        assert (pinned[nbpp:] == 7.0).all()                                          # nothing written past the end
        assert (got["structs"] == want["structs"]).all()
    ref = oracle.fold_batch(bases, offsets, True, False, tt, ct, [2.0], n_threads=8)
    assert_bits_equal(got["bpp"], ref["bpp"], "zero-copy BPP vs oracle")


@pytest.mark.parametrize("contra", [False, True])
def test_fold_sums_and_fold_scores_export(handle, oracle, contra):
    """N3: get_fold_sums(_contra) stand-alone + the FoldScores memo (src/mccaskill_algo.rs:3-22, 213-245, 282-516)
    through rna_fold_sums_batch: every plane bit-equal to the oracle's, in all three kernel routes (shared-memory
    batch, HBM-resident one-CTA, cooperative)."""
    tt, ct, _ = default_tables()
    seqs = load_trnas()[:3] + random_seqs(5, [1, 4, 5, 31, 300]) + [np.random.default_rng(9).integers(0, 4, 450).astype(np.uint8)]
    bases, offsets = pack(seqs)
    got = handle.fold_sums_batch(bases, offsets, contra, False)
    names = handle.SUMS_PLANES
    for s, seq in enumerate(seqs):
        want = oracle.fold_sums(seq, contra, False, tt, ct)
        iu = np.triu_indices(len(seq))
        for p, name in enumerate(names):
            assert_bits_equal(got[s][name][iu], want[p][iu], f"{name} seq {s} (L={len(seq)})")
        assert_bits_equal(got[s]["logz"], want[2][0, len(seq) - 1], "logZ")
    # a lone mid-length sequence takes the cooperative route
    seq = np.random.default_rng(10).integers(0, 4, 500).astype(np.uint8)
    b1, o1 = pack([seq])
    g1 = handle.fold_sums_batch(b1, o1, contra, False)[0]
    want = oracle.fold_sums(seq, contra, False, tt, ct)
    iu = np.triu_indices(500)
    for p, name in enumerate(names):
        assert_bits_equal(g1[name][iu], want[p][iu], f"{name} cooperative route")


@pytest.mark.parametrize("contra", [False, True])
def test_twoloop_scores_export(handle, oracle, contra):
    """N3: FoldScores::twoloop_scores through rna_twoloop_scores — every (i,j,k,l) key in the reference's insertion
    order with its score bit-equal to the oracle's get_2loop_score; sizing call, truncated capacity, short hairpins."""
    import ctypes as C
    tt, ct, _ = default_tables()
    for seq, short in ((load_trnas()[4], False), (random_seqs(8, [37])[0], bool(contra)), (random_seqs(8, [4])[0], False)):
        want = oracle.twoloop_scores(seq, contra, short, tt, ct)
        got = handle.twoloop_scores(seq, contra, short)
        assert got.shape[0] == len(want)
        if want:
            w = np.array([(i, j, k, l) for (i, j, k, l, _) in want], dtype=np.uint16)
            assert (np.stack([got["i"], got["j"], got["k"], got["l"]], axis=1) == w).all(), "keys / order"
            assert_bits_equal(got["score"], np.array([sc for (*_, sc) in want], dtype=np.float32), "two-loop scores")
    # capacity smaller than the count: the count is still reported, the prefix is written
    seq = load_trnas()[0]
    full = handle.twoloop_scores(seq, contra)
    part = np.zeros(10, dtype=handle.TWOLOOP_DTYPE)
    cnt = C.c_uint64(0)
    rc = handle.lib.rna_twoloop_scores(handle.h, seq.ctypes.data_as(C.c_void_p), seq.shape[0], int(contra), 0,
                                       part.ctypes.data_as(C.c_void_p), 10, C.byref(cnt))
    assert rc == 0 and cnt.value == full.shape[0] and (part == full[:10]).all()


def test_multi_device_object_partitions_and_scatters(oracle):
    """rna_multi: LPT partition inside the library, one host thread + handle per device, results scattered into the
    caller's buffers.  Listing device 0 three times exercises partition / scatter on a one-GPU box; on a multi-GPU
    box the default (all visible devices) runs too.  Results are bit-identical to the oracle whatever the partition."""
    import torch
    from rna_algos_b200.api import MultiHandle
    tt, ct, at = default_tables()
    seqs = load_trnas() * 3 + random_seqs(21, [1, 5, 40, 130, 260, 33, 77])
    bases, offsets = pack(seqs)
    gammas = [1.0, 4.0]
    want = oracle.fold_batch(bases, offsets, True, False, tt, ct, gammas, n_threads=8)
    pairs = np.array([(a, b) for a in range(6) for b in range(a + 1, 6)] + [(20, 3), (24, 24)], dtype=np.uint32)
    wantd = oracle.durbin_batch(bases, offsets, pairs, at, n_threads=8)
    configs = [[0, 0, 0]] + ([None] if torch.cuda.device_count() > 1 else [])
    for devs in configs:
        m = MultiHandle(devs, tt, ct, at)
        try:
            got = m.fold_batch(bases, offsets, True, False, gammas)
            assert_bits_equal(got["logz"], want["logz"], "logZ")
            assert_bits_equal(got["bpp"], want["bpp"], "BPP")
            assert_bits_equal(got["expect_acc"], want["expect_acc"], "expect_accuracy")
            assert (got["structs"] == want["structs"]).all()
            busy, units = m.shares()
            assert units.sum() == len(seqs) and (units > 0).all() and (busy > 0).all()
            gd = m.durbin_batch(bases, offsets, pairs)
            assert_bits_equal(gd["probs"], wantd["probs"], "Durbin")
        finally:
            m.close()


def test_sixteen_threads_of_single_calls_are_coalesced(handle, oracle):
    """The reference's call granularity (one sequence per call from the tasks of a thread pool,
    src/bin/centroid_fold.rs:119-132) through rna_queue: 16 threads x single calls are coalesced into batched launches,
    every caller gets its own bit-exact result, and the rate beats the CPU oracle's on as many threads."""
    import threading
    import time
    from rna_algos_b200.api import CallQueue
    tt, ct, _ = default_tables()
    base = load_trnas()
    n_threads, per_thread = 16, 24
    work = [[base[(t + k) % 6] for k in range(per_thread)] for t in range(n_threads)]
    want = {i: oracle.mccaskill(base[i], True, False, tt, ct) for i in range(6)}
    wantc = {i: oracle.centroid(want[i][0], len(base[i]), 2.0) for i in range(6)}
    q = CallQueue(handle)
    errors = []

    def worker(t):
        try:
            for k, seq in enumerate(work[t]):
                i = (t + k) % 6
                bpp, logz, st, ea = q.mccaskill_algo(seq, True, False, centroid_threshold=2.0)
                if not ((bpp.view(np.uint32) == want[i][0].view(np.uint32)).all() and st == wantc[i][0]
                        and np.float32(logz).view(np.uint32) == np.float32(want[i][1]).view(np.uint32)):
                    errors.append((t, k))
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))
    for t in range(2):   # warm-up (allocations, first launches)
        worker(t)
    th = [threading.Thread(target=worker, args=(t,)) for t in range(n_threads)]
    t0 = time.perf_counter()
    for x in th:
        x.start()
    for x in th:
        x.join()
    dt = time.perf_counter() - t0
    stats = q.stats()
    q.close()
    assert not errors, errors[:3]
    rate = n_threads * per_thread / dt
    # the CPU port on the same number of threads, same sequences
    flat = [s for w in work for s in w]
    b, o = pack(flat)
    t0 = time.perf_counter()
    oracle.fold_batch(b, o, True, False, tt, ct, [2.0], n_threads=n_threads)
    cpu_rate = len(flat) / (time.perf_counter() - t0)
    print(f"single calls from {n_threads} threads: {rate:.0f} seq/s in {stats['launches']} launches for {stats['requests']} "
          f"requests; CPU oracle on {n_threads} threads: {cpu_rate:.0f} seq/s")
    assert stats["launches"] < stats["requests"] / 2, stats      # coalescing happened
    assert rate > cpu_rate, (rate, cpu_rate)
