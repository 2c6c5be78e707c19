"""ctypes binding of the CPU oracle (oracle/oracle.c).  Test infrastructure only: imported by tests/,
__graft_entry__.smoke() and bench.py's CPU legs — never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rna_algos_b200.tables import AlignTables, ContraTables, TurnerTables

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

u8p = C.POINTER(C.c_uint8)
u16p = C.POINTER(C.c_uint16)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f32p = C.POINTER(C.c_float)


def build() -> None:
    subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)


def _ptr(a, typ):
    if a is None:
        return None
    return a.ctypes.data_as(typ)


class Oracle:
    def __init__(self, exact: bool = False):
        name = "liboracle_exact.so" if exact else "liboracle.so"
        path = os.path.join(ORACLE_DIR, "_build", name)
        src = os.path.join(ORACLE_DIR, "oracle.c")
        if not os.path.exists(path) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(path)):
            build()
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_ln_exp_1p.restype = C.c_float
        L.orc_ln_exp_1p.argtypes = [C.c_float]
        L.orc_expf.restype = C.c_float
        L.orc_expf.argtypes = [C.c_float]
        L.orc_logsumexp.restype = C.c_float
        L.orc_logsumexp.argtypes = [C.c_float, C.c_float]
        L.orc_mccaskill_algo.restype = C.c_int
        L.orc_mccaskill_algo.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.POINTER(TurnerTables),
                                         C.POINTER(ContraTables), f32p, f32p, f32p, f32p, f32p, f32p]
        L.orc_centroid_fold.restype = C.c_int
        L.orc_centroid_fold.argtypes = [f32p, C.c_int, C.c_float, u8p, u16p, u32p, f32p]
        L.orc_durbin_algo.restype = C.c_int
        L.orc_durbin_algo.argtypes = [u8p, C.c_int, u8p, C.c_int, C.POINTER(AlignTables), f32p]
        L.orc_contra_accumulate.argtypes = [C.POINTER(ContraTables)]
        for fn in ("orc_score_hairpin", "orc_score_multibranch_close", "orc_score_accessible"):
            getattr(L, fn).restype = C.c_double
            getattr(L, fn).argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(TurnerTables),
                                       C.POINTER(ContraTables)]
        L.orc_score_twoloop.restype = C.c_double
        L.orc_score_twoloop.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(TurnerTables), C.POINTER(ContraTables)]
        L.orc_mccaskill_centroid_batch.restype = C.c_int
        L.orc_mccaskill_centroid_batch.argtypes = [u8p, u32p, C.c_uint32, C.c_int, C.c_int,
                                                   C.POINTER(TurnerTables), C.POINTER(ContraTables), f32p,
                                                   C.c_uint32, f32p, f32p, u64p, u8p, f32p, C.c_int, u64p]
        L.orc_durbin_batch.restype = C.c_int
        L.orc_durbin_batch.argtypes = [u8p, u32p, u32p, C.c_uint32, C.POINTER(AlignTables), f32p, u64p,
                                       C.c_int, u64p]
        L.orc_set_inner_threads.argtypes = [C.c_int]
        L.orc_fold_sums.restype = C.c_int
        L.orc_fold_sums.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.POINTER(TurnerTables), C.POINTER(ContraTables), f32p]
        L.orc_is_exact_flavour.restype = C.c_int
        assert bool(L.orc_is_exact_flavour()) == exact

    def set_inner_threads(self, n: int) -> None:
        """Spread the cells of one span of ONE sequence over n host threads (bit-identical; long-sequence checks)."""
        self.lib.orc_set_inner_threads(int(n))

    # ---- single-item calls --------------------------------------------------------------------
    def mccaskill(self, seq: np.ndarray, contra: bool, allows_short: bool, tt, ct, debug: bool = False):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        L = int(seq.shape[0])
        bpp = np.empty(L * (L - 1) // 2, dtype=np.float32)
        logz = C.c_float()
        dbg = [np.empty((L, L), dtype=np.float32) for _ in range(4)] if debug else [None] * 4
        rc = self.lib.orc_mccaskill_algo(_ptr(seq, u8p), L, int(contra), int(allows_short),
                                         C.byref(tt) if tt is not None else None,
                                         C.byref(ct) if ct is not None else None,
                                         _ptr(bpp, f32p), C.byref(logz), *[_ptr(d, f32p) for d in dbg])
        assert rc == 0, rc
        if debug:
            return bpp, float(logz.value), dict(close=dbg[0], external=dbg[1], logprob=dbg[2], m1=dbg[3])
        return bpp, np.float32(logz.value)

    def fold_sums(self, seq: np.ndarray, contra: bool, allows_short: bool, tt, ct) -> np.ndarray:
        """[RNA_SUMS_PLANES][L][L] dense planes of get_fold_sums(_contra) + FoldScores (-inf = key absent)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        L = int(seq.shape[0])
        out = np.empty((10, L, L), dtype=np.float32)
        rc = self.lib.orc_fold_sums(_ptr(seq, u8p), L, int(contra), int(allows_short), C.byref(tt), C.byref(ct),
                                    _ptr(out, f32p))
        assert rc == 0, rc
        return out

    def twoloop_scores(self, seq: np.ndarray, contra: bool, allows_short: bool, tt, ct):
        """FoldScores::twoloop_scores in the reference's insertion order (src/mccaskill_algo.rs:290-324, 395-435):
        list of (i, j, k, l, f32 score).  The visited pairs and the partners' sums_close keys come from fold_sums()."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        L = int(seq.shape[0])
        close = self.fold_sums(seq, contra, allows_short, tt, ct)[0]
        has = close > -np.inf
        canon = {(0, 3), (1, 2), (2, 1), (2, 3), (3, 0), (3, 2)}
        max2 = ct.max_loop_len if contra else tt.max_2loop_len
        minspan = ct.min_span_hairpin_close if contra else tt.min_span_hairpin_close
        sp = _ptr(seq, u8p)
        out = []
        for span in range(2 if (contra and allows_short) else minspan, L + 1):
            for i in range(0, L - span + 1):
                j = i + span - 1
                if (int(seq[i]), int(seq[j])) not in canon:
                    continue
                for k in range(i + 1, j - 1):
                    if k - i - 1 > max2:
                        break
                    for l in range(j - 1, k, -1):
                        if (j - l - 1) + (k - i - 1) > max2:
                            break
                        if has[k, l]:
                            sc = self.lib.orc_score_twoloop(sp, L, i, j, k, l, int(contra), C.byref(tt), C.byref(ct))
                            out.append((i, j, k, l, np.float32(sc)))
        return out

    def centroid(self, bpp: np.ndarray, L: int, gamma: float):
        bpp = np.ascontiguousarray(bpp, dtype=np.float32)
        s = np.empty(L, dtype=np.uint8)
        pairs = np.zeros((max(L, 2), 2), dtype=np.uint16)
        n = C.c_uint32()
        ea = C.c_float()
        rc = self.lib.orc_centroid_fold(_ptr(bpp, f32p), L, C.c_float(gamma), _ptr(s, u8p), _ptr(pairs, u16p),
                                        C.byref(n), C.byref(ea))
        assert rc == 0, rc
        return s.tobytes().decode(), pairs[: n.value].copy(), np.float32(ea.value)

    def durbin(self, sa: np.ndarray, sb: np.ndarray, at):
        sa = np.ascontiguousarray(sa, dtype=np.uint8)
        sb = np.ascontiguousarray(sb, dtype=np.uint8)
        out = np.empty((sa.shape[0] + 2, sb.shape[0] + 2), dtype=np.float32)
        rc = self.lib.orc_durbin_algo(_ptr(sa, u8p), int(sa.shape[0]), _ptr(sb, u8p), int(sb.shape[0]),
                                      C.byref(at), _ptr(out, f32p))
        assert rc == 0, rc
        return out

    # ---- batch calls (threaded; the CPU baseline) ---------------------------------------------
    def fold_batch(self, bases, offsets, contra, allows_short, tt, ct, gammas, n_threads=1,
                   want_bpp=True, inner_threads=1):
        """n_threads: one sequence per task (the CPU baseline's shape).  inner_threads > 1: sequences one after the
        other, the cells of each span spread over that many threads (long sequences; bit-identical)."""
        if inner_threads > 1:
            self.set_inner_threads(inner_threads)
            try:
                return self.fold_batch(bases, offsets, contra, allows_short, tt, ct, gammas, 1, want_bpp, 1)
            finally:
                self.set_inner_threads(1)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        gammas = np.ascontiguousarray(gammas, dtype=np.float32)
        n = offsets.shape[0] - 1
        lens = np.diff(offsets.astype(np.int64))
        bpp_off = np.zeros(n + 1, dtype=np.uint64)
        bpp_off[1:] = np.cumsum(lens * (lens - 1) // 2)
        logz = np.empty(n, dtype=np.float32)
        bpp = np.empty(int(bpp_off[-1]), dtype=np.float32) if want_bpp else None
        ng = gammas.shape[0]
        structs = np.empty((ng, int(offsets[-1])), dtype=np.uint8)
        ea = np.empty((ng, n), dtype=np.float32)
        terms = C.c_uint64()
        rc = self.lib.orc_mccaskill_centroid_batch(
            _ptr(bases, u8p), _ptr(offsets, u32p), n, 1 if contra else 0, int(allows_short),
            C.byref(tt) if tt is not None else None, C.byref(ct) if ct is not None else None,
            _ptr(gammas, f32p), ng, _ptr(logz, f32p), _ptr(bpp, f32p), _ptr(bpp_off, u64p),
            _ptr(structs, u8p), _ptr(ea, f32p), int(n_threads), C.byref(terms))
        assert rc == 0, rc
        return dict(logz=logz, bpp=bpp, bpp_offsets=bpp_off, structs=structs, expect_acc=ea,
                    lse_terms=int(terms.value))

    def durbin_batch(self, bases, offsets, pairs, at, n_threads=1):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        lens = np.diff(offsets.astype(np.int64))
        sizes = (lens[pairs[:, 0]] + 2) * (lens[pairs[:, 1]] + 2)
        po = np.zeros(pairs.shape[0] + 1, dtype=np.uint64)
        po[1:] = np.cumsum(sizes)
        out = np.empty(int(po[-1]), dtype=np.float32)
        terms = C.c_uint64()
        rc = self.lib.orc_durbin_batch(_ptr(bases, u8p), _ptr(offsets, u32p), _ptr(pairs, u32p),
                                       pairs.shape[0], C.byref(at), _ptr(out, f32p), _ptr(po, u64p),
                                       int(n_threads), C.byref(terms))
        assert rc == 0, rc
        return dict(probs=out, prob_offsets=po, lse_terms=int(terms.value))
