"""Host-side mirror of the reference crate's public API for the hot path, over the C ABI.

reference (Rust, heartsh/rna-algos 0.1.37)                     here
-------------------------------------------------------------  ------------------------------------
utils::bytes2seq                     src/utils.rs:562-577       bytes2seq
mccaskill_algo::mccaskill_algo<T>    src/mccaskill_algo.rs:247  mccaskill_algo / Handle.mccaskill_batch
centroid_fold::centroid_fold<T>      src/centroid_fold.rs:25    centroid_fold / Handle.centroid_batch
bin centroid_fold (BPP + MEA sweep)  src/bin/centroid_fold.rs   Handle.fold_batch (fused on device)
durbin_algo::durbin_algo             src/durbin_algo.rs:73      durbin_algo / Handle.durbin_batch

Same argument meaning (``uses_contra_model``, ``allows_short_hairpins``, ``centroid_threshold``) and the
same error behaviour translated to exceptions: the reference panics on a non-ACGU byte or an empty
sequence; here ``RnaError`` is raised.  Everything computes on the GPU through
``librna_algos_b200.so``; there is no CPU path in this package.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import _lib
from . import tables as T

UNPAIR, BASEPAIR_LEFT, BASEPAIR_RIGHT = ".", "(", ")"   # src/utils.rs:123-125
PSEUDO_BASE = 4                                         # src/utils.rs:122
MIN_POW_2, MAX_POW_2 = -7, 10                           # src/bin/centroid_fold.rs:9-10


class RnaError(RuntimeError):
    def __init__(self, code: int, detail: str = ""):
        self.code = code
        super().__init__(f"{_lib.STATUS.get(code, code)}{': ' + detail if detail else ''}")


_CHAR2BASE = np.full(256, 255, dtype=np.uint8)
for _c, _b in (("A", 0), ("C", 1), ("G", 2), ("U", 3)):
    _CHAR2BASE[ord(_c)] = _b
    _CHAR2BASE[ord(_c.lower())] = _b


def bytes2seq(x) -> np.ndarray:
    """ASCII a/c/g/u (either case) -> base codes 0..3; anything else is an error (the reference
    panics: src/utils.rs:570-572; 'T' and 'N' are NOT accepted there either)."""
    if isinstance(x, str):
        x = x.encode()
    a = np.frombuffer(bytes(x), dtype=np.uint8)
    s = _CHAR2BASE[a]
    if (s == 255).any():
        raise RnaError(2, "non-ACGU byte in sequence")
    return s


def read_fasta(path: str) -> List[Tuple[str, np.ndarray]]:
    recs: List[Tuple[str, List[str]]] = []
    with open(path) as fh:
        for line in fh:
            line = line.strip()
            if not line:
                continue
            if line.startswith(">"):
                recs.append((line[1:].split()[0] if len(line) > 1 else "", []))
            elif recs:
                recs[-1][1].append(line)
    return [(rid, bytes2seq("".join(parts))) for rid, parts in recs]


def pack_seqs(seqs: Sequence[np.ndarray]) -> Tuple[np.ndarray, np.ndarray]:
    """Concatenate sequences -> (bases u8, offsets u32[n+1])."""
    lens = np.array([len(s) for s in seqs], dtype=np.int64)
    offsets = np.zeros(len(seqs) + 1, dtype=np.uint32)
    offsets[1:] = np.cumsum(lens)
    bases = np.concatenate([np.asarray(s, dtype=np.uint8) for s in seqs]) if len(seqs) else np.zeros(0, np.uint8)
    return np.ascontiguousarray(bases), offsets


def bpp_offsets_of(offsets: np.ndarray) -> np.ndarray:
    lens = np.diff(offsets.astype(np.int64))
    o = np.zeros(len(lens) + 1, dtype=np.uint64)
    o[1:] = np.cumsum(lens * (lens - 1) // 2)
    return o


def bpp_index(L: int, i: int, j: int) -> int:
    return i * (2 * L - i - 1) // 2 + (j - i - 1)


def sparse_prob_mat(bpp: np.ndarray, L: int) -> Dict[Tuple[int, int], float]:
    """Packed BPP -> the reference's SparseProbMat<T> (HashMap<(i,j), Prob>): one key per closable pair,
    INCLUDING keys whose probability is exactly 0.0 (expf flush), absent keys dropped."""
    out = {}
    for i in range(L):
        base = i * (2 * L - i - 1) // 2
        row = bpp[base: base + L - 1 - i]
        for x in np.nonzero(row != T.BPP_ABSENT)[0]:
            out[(i, i + 1 + int(x))] = float(row[x])
    return out


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def fold_cost(L) -> np.ndarray:
    """Cost model of one McCaskill+centroid unit: cubic split-point loops + the
    O(496 L^2) interior-loop enumeration."""
    L = np.asarray(L, dtype=np.uint64)
    return L * L * L + np.uint64(500) * L * L


def partition_lpt(costs: np.ndarray, n_parts: int) -> np.ndarray:
    """Longest-processing-time-first partition of work units over GPUs (no collective is needed: units
    are independent — src/bin/centroid_fold.rs:119-132 fans them out over threads the same way)."""
    lib = _lib.load()
    costs = np.ascontiguousarray(costs, dtype=np.uint64)
    part = np.zeros(costs.shape[0], dtype=np.uint32)
    rc = lib.rna_partition_lpt(_p(costs), costs.shape[0], n_parts, _p(part))
    if rc:
        raise RnaError(rc)
    return part


class Handle:
    """One GPU.  Owns the device copies of the table blobs and the scratch buffers."""

    def __init__(self, device: int = 0, turner: Optional[T.TurnerTables] = None,
                 contra: Optional[T.ContraTables] = None, align: Optional[T.AlignTables] = None):
        self.lib = _lib.load()
        h = C.c_void_p()
        rc = self.lib.rna_create(device, C.byref(h))
        if rc:
            raise RnaError(rc, "rna_create: a CUDA device is required (no CPU fallback)")
        self.h = h
        self.set_tables(turner, contra, align)

    def set_tables(self, turner=None, contra=None, align=None):
        if turner is not None:
            self._chk(self.lib.rna_set_turner_tables(self.h, C.byref(turner)))
        if contra is not None:
            self._chk(self.lib.rna_set_contra_tables(self.h, C.byref(contra)))
        if align is not None:
            self._chk(self.lib.rna_set_align_tables(self.h, C.byref(align)))

    def set_numeric_mode(self, mode) -> None:
        """"exact" (default: the reference's numerics, bit for bit), "fast" (f32 warp-shuffle reductions, exact math) or
        "fast64" — include/rna_algos_b200.h RNA_NUMERIC_*."""
        code = {"exact": 0, "fast": 1, "fast32": 1, "fast64": 2}.get(mode, mode)
        self._chk(self.lib.rna_set_numeric_mode(self.h, int(code)))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rna_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc: int):
        if rc:
            raise RnaError(rc, (self.lib.rna_last_error(self.h) or b"").decode())

    def stats(self) -> Dict[str, int]:
        s = _lib.CallStats()
        self.lib.rna_get_stats(self.h, C.byref(s))
        return dict(kernel_launches=int(s.kernel_launches), h2d_bytes=int(s.h2d_bytes), d2h_bytes=int(s.d2h_bytes))

    # ---- batched, host buffers ------------------------------------------------------------------
    def fold_batch(self, bases: np.ndarray, offsets: np.ndarray, uses_contra_model: bool,
                   allows_short_hairpins: bool = False, gammas: Iterable[float] = (),
                   want_bpp: bool = True, out: Optional[dict] = None) -> dict:
        """mccaskill_algo for every sequence (+ centroid_fold for every gamma), fused on the GPU.
        `out` may carry preallocated (e.g. pinned) numpy arrays: logz, bpp, structs, expect_acc."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        n = offsets.shape[0] - 1
        g = np.ascontiguousarray(np.array(list(gammas), dtype=np.float32))
        ng = g.shape[0]
        total = int(offsets[-1]) if n else 0
        out = dict(out or {})
        bpp_off = bpp_offsets_of(offsets)
        logz = out.get("logz")
        if logz is None:
            logz = np.empty(n, dtype=np.float32)
        bpp = out.get("bpp") if want_bpp else None
        if want_bpp and bpp is None:
            bpp = np.empty(int(bpp_off[-1]), dtype=np.float32)
        structs = out.get("structs")
        if structs is None:
            structs = np.empty((ng, total), dtype=np.uint8)
        ea = out.get("expect_acc")
        if ea is None:
            ea = np.empty((ng, n), dtype=np.float32)
        self._chk(self.lib.rna_mccaskill_centroid_batch(
            self.h, _p(bases), _p(offsets), n, _lib.MODEL_CONTRA if uses_contra_model else _lib.MODEL_TURNER,
            int(allows_short_hairpins), _p(g) if ng else None, ng, _p(logz), _p(bpp),
            _p(bpp_off) if want_bpp else None, _p(structs) if ng else None, _p(ea) if ng else None))
        return dict(logz=logz, bpp=bpp, bpp_offsets=bpp_off, structs=structs, expect_acc=ea, gammas=g)

    SUMS_PLANES = ("sums_close", "sums_accessible", "sums_external", "sums_rightmost_basepairs_external",
                   "sums_rightmost_basepairs_multibranch", "sums_multibranch", "sums_1ormore_basepairs",
                   "hairpin_scores", "multibranch_close_scores", "accessible_scores")

    def fold_sums_batch(self, bases, offsets, uses_contra_model, allows_short_hairpins=False) -> List[Dict[str, np.ndarray]]:
        """get_fold_sums / get_fold_sums_contra stand-alone + the FoldScores memo (src/mccaskill_algo.rs:3-22, 282-516):
        per sequence a dict of dense L x L arrays (upper triangle filled; -inf = key absent / never written, except
        sums_external whose untouched entries are 0.0 like in the reference) and `logz`."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        n = offsets.shape[0] - 1
        lens = np.diff(offsets.astype(np.int64))
        so = np.zeros(n + 1, dtype=np.uint64)
        so[1:] = np.cumsum(len(self.SUMS_PLANES) * (lens * (lens + 1) // 2))
        flat = np.empty(int(so[-1]), dtype=np.float32)
        logz = np.empty(n, dtype=np.float32)
        self._chk(self.lib.rna_fold_sums_batch(self.h, _p(bases), _p(offsets), n,
                                               _lib.MODEL_CONTRA if uses_contra_model else _lib.MODEL_TURNER,
                                               int(allows_short_hairpins), _p(flat), _p(so), _p(logz)))
        out = []
        for s in range(n):
            L = int(lens[s])
            pl = L * (L + 1) // 2
            iu = np.triu_indices(L)
            d = {"logz": logz[s]}
            for k, name in enumerate(self.SUMS_PLANES):
                m = np.full((L, L), 0.0 if name == "sums_external" else -np.inf, dtype=np.float32)
                m[iu] = flat[int(so[s]) + k * pl: int(so[s]) + (k + 1) * pl]
                d[name] = m
            out.append(d)
        return out

    TWOLOOP_DTYPE = np.dtype([("i", np.uint16), ("j", np.uint16), ("k", np.uint16), ("l", np.uint16), ("score", np.float32)])

    def twoloop_scores(self, seq: np.ndarray, uses_contra_model: bool, allows_short_hairpins: bool = False) -> np.ndarray:
        """FoldScores::twoloop_scores (src/mccaskill_algo.rs:16,320,431) of one sequence: structured array of
        (i, j, k, l, score) in the reference's insertion order."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        model = _lib.MODEL_CONTRA if uses_contra_model else _lib.MODEL_TURNER
        cnt = C.c_uint64(0)
        self._chk(self.lib.rna_twoloop_scores(self.h, _p(seq), seq.shape[0], model, int(allows_short_hairpins), None, 0, C.byref(cnt)))
        out = np.empty(cnt.value, dtype=self.TWOLOOP_DTYPE)
        if cnt.value:
            self._chk(self.lib.rna_twoloop_scores(self.h, _p(seq), seq.shape[0], model, int(allows_short_hairpins),
                                                  _p(out), cnt.value, C.byref(cnt)))
        return out

    def mccaskill_batch(self, bases, offsets, uses_contra_model, allows_short_hairpins=False):
        return self.fold_batch(bases, offsets, uses_contra_model, allows_short_hairpins, gammas=())

    def centroid_batch(self, bpp: np.ndarray, offsets: np.ndarray, gammas: Iterable[float]) -> dict:
        bpp = np.ascontiguousarray(bpp, dtype=np.float32)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        n = offsets.shape[0] - 1
        g = np.ascontiguousarray(np.array(list(gammas), dtype=np.float32))
        structs = np.empty((g.shape[0], int(offsets[-1])), dtype=np.uint8)
        ea = np.empty((g.shape[0], n), dtype=np.float32)
        bpp_off = bpp_offsets_of(offsets)
        self._chk(self.lib.rna_centroid_batch(self.h, _p(bpp), _p(bpp_off), _p(offsets), n, _p(g), g.shape[0],
                                              _p(structs), _p(ea)))
        return dict(structs=structs, expect_acc=ea)

    def durbin_batch(self, bases: np.ndarray, offsets: np.ndarray, pairs: np.ndarray,
                     out: Optional[np.ndarray] = None) -> dict:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        lens = np.diff(offsets.astype(np.int64))
        po = np.zeros(pairs.shape[0] + 1, dtype=np.uint64)
        po[1:] = np.cumsum((lens[pairs[:, 0]] + 2) * (lens[pairs[:, 1]] + 2))
        if out is None:
            out = np.empty(int(po[-1]), dtype=np.float32)
        self._chk(self.lib.rna_durbin_batch(self.h, _p(bases), _p(offsets), offsets.shape[0] - 1, _p(pairs),
                                            pairs.shape[0], _p(out), _p(po)))
        return dict(probs=out, prob_offsets=po)

    # ---- single item (the reference's call granularity) ----------------------------------------
    def mccaskill_algo(self, seq: np.ndarray, uses_contra_model: bool, allows_short_hairpins: bool = False):
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        L = seq.shape[0]
        if L == 0:
            raise RnaError(3, "empty sequence")
        bpp = np.empty(L * (L - 1) // 2, dtype=np.float32)
        logz = C.c_float()
        self._chk(self.lib.rna_mccaskill_algo(self.h, _p(seq), L, int(uses_contra_model),
                                              int(allows_short_hairpins), _p(bpp), C.byref(logz)))
        return bpp, np.float32(logz.value)

    def centroid_fold(self, bpp: np.ndarray, seq_len: int, centroid_threshold: float):
        bpp = np.ascontiguousarray(bpp, dtype=np.float32)
        s = np.empty(seq_len, dtype=np.uint8)
        pairs = np.zeros((max(seq_len, 2), 2), dtype=np.uint16)
        npairs = C.c_uint32()
        ea = C.c_float()
        self._chk(self.lib.rna_centroid_fold(self.h, _p(bpp), seq_len, C.c_float(centroid_threshold), _p(s),
                                             _p(pairs), C.byref(npairs), C.byref(ea)))
        return s.tobytes().decode(), pairs[: npairs.value].copy(), np.float32(ea.value)

    def durbin_algo(self, seq_a: np.ndarray, seq_b: np.ndarray) -> np.ndarray:
        a = np.ascontiguousarray(seq_a, dtype=np.uint8)
        b = np.ascontiguousarray(seq_b, dtype=np.uint8)
        out = np.empty((a.shape[0] + 2, b.shape[0] + 2), dtype=np.float32)
        self._chk(self.lib.rna_durbin_algo(self.h, _p(a), a.shape[0], _p(b), b.shape[0], _p(out)))
        return out


class CallQueue:
    """Thread-safe front end of a Handle for the reference's call granularity (one sequence per call from a thread
    pool, src/bin/centroid_fold.rs:119-132): concurrent calls are coalesced into batched launches (rna_queue)."""

    def __init__(self, handle: Handle):
        self.handle = handle
        self.lib = handle.lib
        q = C.c_void_p()
        rc = self.lib.rna_queue_create(handle.h, C.byref(q))
        if rc:
            raise RnaError(rc)
        self.q = q

    def mccaskill_algo(self, seq, uses_contra_model: bool, allows_short_hairpins: bool = False,
                       centroid_threshold: Optional[float] = None):
        """-> (packed bpp, logz) or, with a threshold, (packed bpp, logz, dot-bracket string, expect_accuracy)."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        L = seq.shape[0]
        bpp = np.empty(L * (L - 1) // 2, dtype=np.float32)
        logz = C.c_float()
        st = np.empty(L, dtype=np.uint8) if centroid_threshold is not None else None
        ea = C.c_float()
        rc = self.lib.rna_queue_mccaskill_algo(self.q, _p(seq), L, int(uses_contra_model), int(allows_short_hairpins), _p(bpp),
                                               C.byref(logz), C.c_float(centroid_threshold or 0.0), _p(st),
                                               C.byref(ea) if st is not None else None)
        if rc:
            raise RnaError(rc, (self.lib.rna_last_error(self.handle.h) or b"").decode())
        if st is None:
            return bpp, np.float32(logz.value)
        return bpp, np.float32(logz.value), st.tobytes().decode(), np.float32(ea.value)

    def stats(self):
        r, l = C.c_uint64(), C.c_uint64()
        self.lib.rna_queue_stats(self.q, C.byref(r), C.byref(l))
        return dict(requests=int(r.value), launches=int(l.value))

    def close(self):
        if getattr(self, "q", None):
            self.lib.rna_queue_destroy(self.q)
            self.q = None


class MultiHandle:
    """Every GPU of the box behind one object (include/rna_algos_b200.h rna_multi): the units of a call are LPT-
    partitioned over the devices inside the library, one host thread per device, no collective."""

    def __init__(self, devices: Optional[Sequence[int]] = None, turner=None, contra=None, align=None):
        self.lib = _lib.load()
        m = C.c_void_p()
        devs = np.ascontiguousarray(np.array(list(devices or []), dtype=np.int32))
        rc = self.lib.rna_multi_create(_p(devs) if len(devs) else None, len(devs), C.byref(m))
        if rc:
            raise RnaError(rc, "rna_multi_create: CUDA devices are required (no CPU fallback)")
        self.m = m
        if turner is not None:
            self._chk(self.lib.rna_multi_set_turner_tables(self.m, C.byref(turner)))
        if contra is not None:
            self._chk(self.lib.rna_multi_set_contra_tables(self.m, C.byref(contra)))
        if align is not None:
            self._chk(self.lib.rna_multi_set_align_tables(self.m, C.byref(align)))

    def _chk(self, rc: int):
        if rc:
            raise RnaError(rc, (self.lib.rna_multi_last_error(self.m) or b"").decode())

    @property
    def num_devices(self) -> int:
        return int(self.lib.rna_multi_num_devices(self.m))

    def close(self):
        if getattr(self, "m", None):
            self.lib.rna_multi_destroy(self.m)
            self.m = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def shares(self):
        nd = self.num_devices
        busy = np.zeros(nd, dtype=np.float64)
        units = np.zeros(nd, dtype=np.uint32)
        self.lib.rna_multi_last_shares(self.m, _p(busy), _p(units))
        return busy, units

    def fold_batch(self, bases, offsets, uses_contra_model, allows_short_hairpins=False, gammas=(), want_bpp=True) -> dict:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        n = offsets.shape[0] - 1
        g = np.ascontiguousarray(np.array(list(gammas), dtype=np.float32))
        ng = g.shape[0]
        total = int(offsets[-1]) if n else 0
        bpp_off = bpp_offsets_of(offsets)
        logz = np.empty(n, dtype=np.float32)
        bpp = np.empty(int(bpp_off[-1]), dtype=np.float32) if want_bpp else None
        structs = np.empty((ng, total), dtype=np.uint8)
        ea = np.empty((ng, n), dtype=np.float32)
        self._chk(self.lib.rna_multi_mccaskill_centroid_batch(
            self.m, _p(bases), _p(offsets), n, _lib.MODEL_CONTRA if uses_contra_model else _lib.MODEL_TURNER,
            int(allows_short_hairpins), _p(g) if ng else None, ng, _p(logz), _p(bpp),
            _p(bpp_off) if want_bpp else None, _p(structs) if ng else None, _p(ea) if ng else None))
        return dict(logz=logz, bpp=bpp, bpp_offsets=bpp_off, structs=structs, expect_acc=ea, gammas=g)

    def durbin_batch(self, bases, offsets, pairs) -> dict:
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint32)
        pairs = np.ascontiguousarray(pairs, dtype=np.uint32).reshape(-1, 2)
        lens = np.diff(offsets.astype(np.int64))
        po = np.zeros(pairs.shape[0] + 1, dtype=np.uint64)
        po[1:] = np.cumsum((lens[pairs[:, 0]] + 2) * (lens[pairs[:, 1]] + 2))
        out = np.empty(int(po[-1]), dtype=np.float32)
        self._chk(self.lib.rna_multi_durbin_batch(self.m, _p(bases), _p(offsets), offsets.shape[0] - 1, _p(pairs),
                                                  pairs.shape[0], _p(out), _p(po)))
        return dict(probs=out, prob_offsets=po)


_default: Optional[Handle] = None
_by_tables: Dict[bytes, Handle] = {}


def default_handle() -> Handle:
    """The handle behind the module-level functions.  Score tables: the genuine rna-ss-params blobs
    (`turner2004.tbl`, `contrafold_v202.tbl`, written by tools/ref_dump) from the directory named by
    $RNA_ALGOS_B200_TABLES.  The in-repo STAND-IN tables are not the reference's numbers: they are used only when
    $RNA_ALGOS_B200_ALLOW_STANDIN=1 opts in, with a warning; otherwise this raises RNA_ERR_NO_TABLES."""
    global _default
    if _default is None:
        import os
        import warnings
        lib = _lib.load()
        d = os.environ.get("RNA_ALGOS_B200_TABLES")
        if d:
            tt, ct = T.load_genuine_tables(d)
        elif os.environ.get("RNA_ALGOS_B200_ALLOW_STANDIN") == "1":
            warnings.warn("rna_algos_b200: using the STAND-IN Turner / CONTRAfold tables; results differ from the "
                          "reference's (set RNA_ALGOS_B200_TABLES to a directory with the genuine blobs)", stacklevel=2)
            tt, ct = T.standin_turner_tables(), T.standin_contra_tables(lib)
        else:
            raise RnaError(5, "no score tables: set RNA_ALGOS_B200_TABLES to a directory holding turner2004.tbl and "
                              "contrafold_v202.tbl (tools/ref_dump writes them from rna-ss-params), pass a Handle, or "
                              "opt in to the stand-in tables with RNA_ALGOS_B200_ALLOW_STANDIN=1")
        _default = Handle(0, tt, ct, T.contralign_tables())
    return _default


def _handle_for(handle: Optional[Handle], contra: Optional[T.ContraTables] = None,
                align: Optional[T.AlignTables] = None) -> Handle:
    """The reference takes `&FoldScoreSets` / `&AlignScores` PER CALL (src/mccaskill_algo.rs:247-255,
    src/durbin_algo.rs:73).  A call that passes its own tables gets a handle of its own, cached by the blob's bytes,
    so the tables of the shared default handle are never changed behind other callers' backs."""
    if handle is not None:
        if contra is not None or align is not None:
            handle.set_tables(contra=contra, align=align)
        return handle
    if contra is None and align is None:
        return default_handle()
    key = (bytes(contra) if contra is not None else b"") + b"|" + (bytes(align) if align is not None else b"")
    h = _by_tables.get(key)
    if h is None:
        if len(_by_tables) >= 8:   # bounded: drop the oldest
            _by_tables.pop(next(iter(_by_tables))).close()
        h = Handle(0, None, contra, align if align is not None else T.contralign_tables())
        _by_tables[key] = h
    return h


def mccaskill_algo(seq, uses_contra_model: bool, allows_short_hairpins: bool = False,
                   fold_score_sets: Optional[T.ContraTables] = None, handle: Optional[Handle] = None):
    """mccaskill_algo(seq, uses_contra_model, allows_short_hairpins, &fold_score_sets) ->
    SparseProbMat (as a dict).  `fold_score_sets` is read only by the CONTRAfold model, as in the reference; the Turner
    model reads the handle's Turner blob (rna-ss-params consts in the reference)."""
    if uses_contra_model and fold_score_sets is not None:
        h = _handle_for(handle, contra=fold_score_sets)
    else:
        h = handle or default_handle()
    bpp, _ = h.mccaskill_algo(seq, uses_contra_model, allows_short_hairpins)
    return sparse_prob_mat(bpp, len(seq))


def centroid_fold(basepair_probs, seq_len: int, centroid_threshold: float, handle: Optional[Handle] = None):
    """centroid_fold(&basepair_probs, seq_len, centroid_threshold) -> (basepair_pos_pairs, expect_accuracy).
    `basepair_probs` is a packed array or the dict form returned by mccaskill_algo."""
    h = handle or default_handle()
    if isinstance(basepair_probs, dict):
        bpp = np.full(seq_len * (seq_len - 1) // 2, T.BPP_ABSENT, dtype=np.float32)
        for (i, j), p in basepair_probs.items():
            bpp[bpp_index(seq_len, i, j)] = p
    else:
        bpp = basepair_probs
    _, pairs, ea = h.centroid_fold(bpp, seq_len, centroid_threshold)
    return [tuple(int(v) for v in p) for p in pairs], float(ea)


def get_fold_str(pairs, seq_len: int) -> str:
    """src/bin/centroid_fold.rs:197-207"""
    s = [UNPAIR] * seq_len
    for i, j in pairs:
        s[int(i)] = BASEPAIR_LEFT
        s[int(j)] = BASEPAIR_RIGHT
    return "".join(s)


def durbin_algo(seq_pair, align_scores: Optional[T.AlignTables] = None, handle: Optional[Handle] = None):
    """durbin_algo(&(seq_a, seq_b), &align_scores) -> ProbMat.  The reference's callers pad both
    sequences with PSEUDO_BASE first (src/bin/durbin_algo.rs:48-50); pass them either way — sentinels
    are stripped and re-added by the library, and the result is the same sentinel-indexed matrix."""
    h = _handle_for(handle, align=align_scores)
    a, b = (np.asarray(s, dtype=np.uint8) for s in seq_pair)
    if len(a) >= 2 and a[0] == PSEUDO_BASE and a[-1] == PSEUDO_BASE:
        a = a[1:-1]
    if len(b) >= 2 and b[0] == PSEUDO_BASE and b[-1] == PSEUDO_BASE:
        b = b[1:-1]
    return h.durbin_algo(a, b)


def centroid_threshold_sweep() -> List[float]:
    """The default threshold range of the reference's binary: 2^-7 .. 2^10 (src/bin/centroid_fold.rs:148-149)."""
    return [float(np.float32(2.0) ** p) for p in range(MIN_POW_2, MAX_POW_2 + 1)]
