"""Shared helpers for the test-suite."""
from __future__ import annotations

import os

import numpy as np

from rna_algos_b200 import tables as T

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRNA_FASTA = os.path.join(ROOT, "tests", "golden", "sampled_trnas.fa")
_CODE = {"A": 0, "C": 1, "G": 2, "U": 3}


def load_trnas():
    """The 6 tRNAs of the reference's assets/sampled_trnas.fa (L = 84, 74, 73, 73, 68, 89)."""
    seqs, cur = [], None
    for line in open(TRNA_FASTA).read().splitlines():
        line = line.strip()
        if line.startswith(">"):
            seqs.append([])
        elif line:
            seqs[-1].append(line)
    return [np.array([_CODE[c] for c in "".join(p)], dtype=np.uint8) for p in seqs]


def random_seqs(seed: int, lengths):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 4, size=int(L)).astype(np.uint8) for L in lengths]


def pack(seqs):
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum(lens)
    return np.concatenate(seqs).astype(np.uint8), offsets


_tables_cache = {}


def default_tables():
    if "d" not in _tables_cache:
        _tables_cache["d"] = (T.standin_turner_tables(), T.standin_contra_tables(), T.contralign_tables())
    return _tables_cache["d"]


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def assert_bits_equal(got, want, what=""):
    g, w = bits(got), bits(want)
    assert g.shape == w.shape, (what, g.shape, w.shape)
    bad = np.nonzero(g != w)[0]
    if bad.size:
        k = int(bad[0])
        gf = np.ascontiguousarray(got, dtype=np.float32).ravel()
        wf = np.ascontiguousarray(want, dtype=np.float32).ravel()
        raise AssertionError(f"{what}: {bad.size} of {g.size} values differ bitwise; first at {k}: got {gf[k]!r} want {wf[k]!r}")
