"""The v2 kernel's parallel decomposition (roles X/Y/Z pipelined over anti-diagonals, fold_phases.cuh),
compiled for the HOST and run barrier by barrier (tests/emu), must be bit-identical to the oracle for any
lane count and any execution order of lanes and roles between two barriers.  No GPU needed."""
import numpy as np
import pytest

from common import default_tables, load_trnas, random_seqs
from emu_lib import Emu
from oracle_lib import Oracle
from rna_algos_b200 import tables as T


@pytest.fixture(scope="module")
def emu():
    return Emu()


@pytest.fixture(scope="module")
def oracle():
    return Oracle()


def check(emu, oracle, seqs, contra, short, tt, ct, **kw):
    for s in seqs:
        wb, wz = oracle.mccaskill(s, contra, short, tt, ct)
        gb, gz = emu.fold(s, contra, short, tt, ct, **kw)
        assert (gb.view(np.uint32) == wb.view(np.uint32)).all(), (len(s), contra, kw)
        assert np.float32(gz).view(np.uint32) == np.float32(wz).view(np.uint32), (len(s), contra, kw)


@pytest.mark.parametrize("contra", [False, True])
def test_trnas(emu, oracle, contra):
    tt, ct, _ = default_tables()
    check(emu, oracle, load_trnas(), contra, False, tt, ct)
    check(emu, oracle, load_trnas()[:3], contra, False, tt, ct, order=1, nX=7, nY=5, nZ=3)


@pytest.mark.parametrize("contra", [False, True])
def test_edge_lengths(emu, oracle, contra):
    tt, ct, _ = default_tables()
    check(emu, oracle, random_seqs(11, list(range(1, 41))), contra, False, tt, ct, order=1)


def test_allows_short(emu, oracle):
    tt, ct, _ = default_tables()
    check(emu, oracle, random_seqs(12, [2, 3, 4, 5, 9, 17, 33, 60]), True, True, tt, ct)


@pytest.mark.parametrize("seed", [1, 2])
@pytest.mark.parametrize("contra", [False, True])
def test_random_tables(emu, oracle, seed, contra):
    rt, rc = T.random_turner_tables(100 + seed), T.random_contra_tables(200 + seed)
    seqs = random_seqs(20 + seed, [5, 8, 13, 21, 34, 55, 76, 100])
    rng = np.random.default_rng(seed)
    seqs += [rng.choice(4, size=64, p=[0.15, 0.35, 0.35, 0.15]).astype(np.uint8) for _ in range(3)]
    check(emu, oracle, seqs, contra, False, rt, rc, order=seed % 2)


@pytest.mark.parametrize("contra", [False, True])
def test_on_the_fly_fallback(emu, oracle, contra):
    """tcap = 0: no term streams (two-loop scores computed inside the chains); tcap = 1000: some sequences
    fit their stream slot, some fall back."""
    tt, ct, _ = default_tables()
    check(emu, oracle, load_trnas()[:2] + random_seqs(5, [12, 40]), contra, False, tt, ct, tcap=0)
    check(emu, oracle, random_seqs(6, [20, 30, 60, 90]), contra, False, tt, ct, tcap=1000, order=1)


@pytest.mark.parametrize("contra", [False, True])
def test_one_diagonal_schedule(emu, oracle, contra):
    """order = 2: one diagonal per step with whole folds (the cooperative long-sequence kernel's schedule), with and
    without term streams."""
    tt, ct, _ = default_tables()
    seqs = load_trnas()[:2] + random_seqs(9, [1, 4, 5, 6, 17, 40, 131])
    check(emu, oracle, seqs, contra, False, tt, ct, order=2)
    check(emu, oracle, seqs[:4], contra, False, tt, ct, order=2, tcap=0, nX=5, nY=3, nZ=2)
    rt, rc = T.random_turner_tables(301), T.random_contra_tables(302)
    check(emu, oracle, random_seqs(10, [33, 64, 90]), contra, False, rt, rc, order=2)


def test_mid_length(emu, oracle):
    tt, ct, _ = default_tables()
    check(emu, oracle, random_seqs(31, [150, 260]), True, False, tt, ct, nX=64, nY=128, nZ=128)
    check(emu, oracle, random_seqs(32, [201]), False, False, tt, ct, order=1)


@pytest.mark.parametrize("contra", [False, True])
def test_pair_split_schedule(emu, oracle, contra):
    """order = 3: the cooperative kernel's inside pass — pair steps, the three dense chains of a cell on three
    lanes, chains of two diagonals side by side, R/Rm and sums_1ormore finished in a second phase."""
    tt, ct, _ = default_tables()
    seqs = load_trnas()[:2] + random_seqs(9, [1, 2, 3, 4, 5, 6, 7, 17, 40, 131])
    check(emu, oracle, seqs, contra, False, tt, ct, order=3)
    check(emu, oracle, seqs[:6], contra, False, tt, ct, order=3, tcap=0, nX=5, nY=3, nZ=40)
    rt, rc = T.random_turner_tables(311), T.random_contra_tables(312)
    check(emu, oracle, random_seqs(10, [33, 64, 90]), contra, False, rt, rc, order=3, nZ=96)
    if contra:
        check(emu, oracle, random_seqs(13, [2, 3, 5, 9, 33]), True, True, tt, ct, order=3)


def test_phases_are_sanitizer_clean(tmp_path):
    """compute-sanitizer is not available on the GPU pool; the same phase code runs here on exactly-sized host
    buffers under AddressSanitizer + UndefinedBehaviorSanitizer instead (all schedules, both models)."""
    import os
    import subprocess
    import sys
    from emu_lib import SRC
    from common import ROOT
    asan = subprocess.run(["g++", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not asan or not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("libasan not available")
    lib = str(tmp_path / "_emu_asan.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-fsanitize=address,undefined",
                    "-fno-sanitize-recover=undefined", "-fno-omit-frame-pointer", "-shared", "-fPIC", "-o", lib, SRC], check=True, capture_output=True)
    env = dict(os.environ, LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "asan_run.py"), ROOT, lib], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "asan-clean" in r.stdout, r.stderr[-3000:]
