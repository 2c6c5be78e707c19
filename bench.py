#!/usr/bin/env python
"""bench.py — McCaskill BPP + gamma-centroid throughput (sequences/s, DP cells/s) on N B200s.

Contract (one JSON line on rank 0):
  python bench.py --gpus N --steps K --warmup W            this repo's CUDA path
  python bench.py --impl reference --gpus N ...            the reference algorithm's CPU path (oracle port,
                                                           all host cores, bounded sample per step)

A "step" is one pass of the hot path (rna_mccaskill_centroid_batch*: inside, outside, BPP, centroid
fold + traceback for every sequence) over one batch.  Default workload = BASELINE.json configs[1]: the
6 tRNAs of assets/sampled_trnas.fa (copied to tests/golden/) tiled to --nseq sequences per GPU, CONTRAfold
v2.02 model, centroid threshold gamma = 1.  Per-GPU work is fixed as N grows ("weak" scaling); ranks are
independent (no collective on the data path).

  value      sequences/s with inputs and outputs resident in HBM (device pointers, *_dev entry point),
             CUDA events on the launching stream, max over ranks.
  e2e        sequences/s through the host-buffer C-ABI call (pinned host memory in, pinned host memory out;
             H2D + kernels + D2H inside the timed region).
  roofline   the roof that binds this path: FP32 NON-FMA issue (counted logsumexp terms x 14 FP32 instructions / s against
             the rate a micro-kernel reaches in the same process, tools/peak_microbench.cu); the HBM view (algorithmic
             bytes / time / MEASURED_PEAKS.json hbm_gbs) rides along under "hbm".
  cpu_baseline  the oracle port (tests-only code) on all host cores over a bounded sample.
  configs    (N = 1) short legs for every BASELINE.json config, each with a bit-parity spot check against the oracle:
             Turner tRNAs, gamma = 2 and the 18-threshold sweep, the seeded Rfam-like families, single long sequences
             (1k / 2k / 4k nt, both models, reference-exact and FAST), Durbin on the intra-family pairs.
  strong     (N > 1) strong scaling of the fixed Rfam-like global batch, LPT-partitioned, with per-rank busy time.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "mccaskill_bpp_centroid_sequences_per_s"
UNIT = "sequences/s"
FP32_INSTR_PER_LSE = 14          # FP32-pipe instructions per reference-exact LSE term


# ----------------------------------------------------------------------------------------------------
# workloads (synthetic / bundled; nothing is read from /root/reference)
# ----------------------------------------------------------------------------------------------------
def make_workload(name: str, nseq: int, seed: int = 20251018, length: int = 76):
    from common import load_trnas
    rng = np.random.default_rng(seed)
    if name in ("trna_contra", "trna_turner"):
        base = load_trnas()
        seqs = [base[i % 6] for i in range(nseq)]
        contra = name == "trna_contra"
        desc = (f"6 tRNAs of assets/sampled_trnas.fa (L=84,74,73,73,68,89) tiled to {nseq} seqs/GPU, "
                f"{'CONTRAfold v2.02' if contra else 'Turner 2004'}, BPP + centroid gamma=1")
    elif name in ("random76_contra", "random76_turner"):
        seqs = [rng.integers(0, 4, size=length).astype(np.uint8) for _ in range(nseq)]
        contra = name.endswith("contra")
        desc = f"{nseq} i.i.d. uniform ACGU sequences of L={length} per GPU, {'CONTRAfold' if contra else 'Turner'}"
    elif name in ("rfam_synth_contra", "rfam_synth_turner"):
        # stand-in for the missing Rfam seed file: log-uniform lengths in [50, 500]
        lens = np.exp(rng.uniform(np.log(50), np.log(500), size=nseq)).astype(int)
        seqs = [rng.integers(0, 4, size=int(L)).astype(np.uint8) for L in lens]
        contra = name.endswith("contra")
        desc = f"{nseq} synthetic Rfam-like sequences (log-uniform 50..500 nt) per GPU"
    elif name in ("long_turner", "long_contra"):
        seqs = [rng.integers(0, 4, size=length).astype(np.uint8) for _ in range(nseq)]
        contra = name.endswith("contra")
        desc = f"{nseq} random sequence(s) of L={length}, HBM-resident multi-CTA wavefront"
    else:
        raise SystemExit(f"unknown workload {name}")
    return seqs, contra, desc


def make_families(n_families: int, seed: int = 20251019):
    """Stand-in for assets/rfam_seed_stas_v14.3.sth (missing from the reference tree): SURVEY.md §8(d) config 3 —
    per family a random parent of length ~ log-uniform[50, 500] and 2-10 members derived from it by 15 %
    substitutions + 5 % indels (members of a family have similar lengths)."""
    rng = np.random.default_rng(seed)
    fams = []
    for _ in range(n_families):
        L = int(np.exp(rng.uniform(np.log(50), np.log(500))))
        parent = rng.integers(0, 4, size=L).astype(np.uint8)
        members = []
        for _m in range(int(rng.integers(2, 11))):
            out = []
            for b in parent:
                r = rng.random()
                if r < 0.025:
                    continue                                   # deletion
                if r < 0.05:
                    out.append(int(rng.integers(0, 4)))        # insertion before the base
                if rng.random() < 0.15:
                    b = (int(b) + int(rng.integers(1, 4))) % 4  # substitution by a different base
                out.append(int(b))
            members.append(np.array(out if out else [0], dtype=np.uint8))
        fams.append(members)
    return fams


def family_batch(n_seqs_target: int):
    """Sequences of the first families up to ~n_seqs_target, and all intra-family pairs (a < b) among them."""
    seqs, pairs = [], []
    for members in make_families(max(8, n_seqs_target // 4)):
        if len(seqs) >= n_seqs_target:
            break
        base = len(seqs)
        seqs.extend(members)
        pairs.extend((base + a, base + b) for a in range(len(members)) for b in range(a + 1, len(members)))
    return seqs, np.array(pairs, dtype=np.uint32).reshape(-1, 2)


def cells_of(lens: np.ndarray) -> int:
    """DP cells of McCaskill+centroid = 3 triangular passes (inside, outside, MEA): SURVEY.md §8(d)."""
    lens = lens.astype(np.int64)
    return int((3 * lens * (lens + 1) // 2).sum())


def algorithmic_bytes(lens: np.ndarray, n_gammas: int) -> int:
    """Compulsory HBM traffic of one step: read bases + offsets, write packed BPP, logZ, structures, E[acc]."""
    lens = lens.astype(np.int64)
    return int((lens + 4 + 8 + 4 * lens * (lens - 1) // 2 + 4 + n_gammas * (lens + 4)).sum())


# ----------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------------
# CPU leg: the oracle port on all host cores over a bounded sample
# ----------------------------------------------------------------------------------------------------
def cpu_run(seqs, contra, gammas, n_threads):
    from common import default_tables, pack
    from oracle_lib import Oracle
    tt, ct, _ = default_tables()
    orc = cpu_run.orc = getattr(cpu_run, "orc", None) or Oracle()
    bases, offsets = pack(seqs)
    t0 = time.perf_counter()
    r = orc.fold_batch(bases, offsets, contra, False, tt, ct, gammas, n_threads=n_threads)
    return time.perf_counter() - t0, r


def cpu_sample(seqs_all, contra, gammas, n_threads, target_s):
    """Pick a prefix of the workload that takes about target_s on the host, run it, return (rate, n, dt, lse/seq)."""
    k = min(len(seqs_all), 6 * max(1, n_threads))
    while True:   # grow the probe until it is long enough to give a stable rate
        probe = seqs_all[:k]
        dt, r = cpu_run(probe, contra, gammas, n_threads)
        if dt >= min(0.5, 0.25 * target_s) or k >= len(seqs_all):
            break
        k = min(len(seqs_all), 2 * k)
    rate = len(probe) / dt
    n = int(max(len(probe), min(len(seqs_all), rate * target_s)))
    n = max(6, n // 6 * 6) if len(seqs_all) >= 6 else n
    dt, r = cpu_run(seqs_all[:n], contra, gammas, n_threads)
    return n / dt, n, dt, r["lse_terms"] / n


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    seqs, contra, desc = make_workload(args.workload, args.nseq, length=args.length)
    gammas = [float(g) for g in args.gammas.split(",")]
    cores = os.cpu_count() or 1
    lens = np.array([len(s) for s in seqs])
    # bounded sample per step: ~args.ref_step_seconds of host work
    rate, n, _, _ = cpu_sample(seqs, contra, gammas, cores, args.ref_step_seconds)
    sample = seqs[:n]
    for _ in range(args.warmup):
        cpu_run(sample, contra, gammas, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(sample, contra, gammas, cores)
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    slens = lens[:n]
    out = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (bundled tRNAs tiled)",
        "config": {"workload": desc, "gammas": gammas, "nseq_per_gpu": args.nseq},
        "cells_per_s": cells_of(slens) * args.steps / dt,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"first {n} sequences of the workload per step (oracle/oracle.c, one sequence per "
                                   f"task on {cores} threads, like benches/benches.rs:24-41)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------
# measured peaks of this process's GPU (tools/peak_microbench.cu): FP32 non-FMA issue, MUFU, shared memory
# ----------------------------------------------------------------------------------------------------
def measure_peaks(device: int):
    path = os.path.join(ROOT, "tools", "_build", "libpeaks.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    lib.peaks_measure.restype = C.c_int
    lib.peaks_measure.argtypes = [C.c_int, C.POINTER(C.c_double)]
    out = (C.c_double * 4)()
    if lib.peaks_measure(device, out) != 0:
        return None
    return {"fp32_nofma_ginstr_s": out[0] / 1e9, "mufu_ginstr_s": out[1] / 1e9, "smem_gbs": out[2] / 1e9, "sms": int(out[3])}


# ----------------------------------------------------------------------------------------------------
# a batch resident in HBM + the *_dev entry point, timed with CUDA events on the launching stream
# ----------------------------------------------------------------------------------------------------
class FoldRunner:
    def __init__(self, h, seqs, contra, gammas, dev, stream):
        import torch
        from common import pack
        from rna_algos_b200 import _lib
        from rna_algos_b200.api import bpp_offsets_of
        self.h, self.stream, self.torch = h, stream, torch
        self.bases, self.offsets = pack(seqs)
        self.n = len(seqs)
        self.lens = np.diff(self.offsets.astype(np.int64))
        self.total = int(self.offsets[-1])
        self.bpp_off = bpp_offsets_of(self.offsets)
        self.bpp_total = int(self.bpp_off[-1])
        self.ng = len(gammas)
        t = self.t = {}
        t["bases"] = torch.from_numpy(self.bases).to(dev)
        t["offsets"] = torch.from_numpy(self.offsets.view(np.int32)).to(dev)
        t["bppoff"] = torch.from_numpy(self.bpp_off.view(np.int64)).to(dev)
        t["gammas"] = torch.tensor(list(gammas) or [1.0], dtype=torch.float32, device=dev)
        t["logz"] = torch.empty(self.n, dtype=torch.float32, device=dev)
        t["bpp"] = torch.empty(max(self.bpp_total, 1), dtype=torch.float32, device=dev)
        t["structs"] = torch.empty(max(self.ng * self.total, 1), dtype=torch.uint8, device=dev)
        t["ea"] = torch.empty(max(self.ng * self.n, 1), dtype=torch.float32, device=dev)
        fb = self.fb = _lib.FoldBatchDev()
        fb.h_offsets = self.offsets.ctypes.data
        fb.d_bases = t["bases"].data_ptr(); fb.d_offsets = t["offsets"].data_ptr(); fb.d_bpp_offsets = t["bppoff"].data_ptr()
        fb.n_seqs = self.n; fb.total_len = self.total; fb.max_len = int(self.lens.max())
        fb.model = _lib.MODEL_CONTRA if contra else _lib.MODEL_TURNER
        fb.allows_short_hairpins = 0
        fb.d_gammas = t["gammas"].data_ptr(); fb.n_gammas = self.ng
        fb.d_out_logz = t["logz"].data_ptr(); fb.d_out_bpp = t["bpp"].data_ptr()
        fb.d_out_structs = t["structs"].data_ptr(); fb.d_out_expect_acc = t["ea"].data_ptr()
        self.sptr = C.c_void_p(stream.cuda_stream)

    def step(self):
        lib = self.h.lib
        rc = lib.rna_mccaskill_centroid_batch_dev(self.h.h, C.byref(self.fb), self.sptr)
        if rc:
            raise RuntimeError(f"rna_mccaskill_centroid_batch_dev rc={rc}: {lib.rna_last_error(self.h.h)}")

    def timed(self, steps, warmup):
        torch = self.torch
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(self.stream)
        for _ in range(steps):
            self.step()
        ev1.record(self.stream)
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1) / steps

    def results(self, idx):
        """(logz, packed bpp list, structs [ng][...]) of the sequences `idx`, copied back from HBM."""
        logz = self.t["logz"].cpu().numpy()
        bpp = self.t["bpp"].cpu().numpy()
        st = self.t["structs"].cpu().numpy()[: self.ng * self.total].reshape(self.ng, self.total) if self.ng else None
        out = []
        for s in idx:
            lo, hi = int(self.bpp_off[s]), int(self.bpp_off[s + 1])
            out.append((logz[s], bpp[lo:hi], None if st is None else st[:, self.offsets[s]:self.offsets[s + 1]]))
        return out


def parity_sample(runner, seqs, contra, gammas, idx, inner_threads=1):
    """Bit-compare the device-resident results of the sequences `idx` with the oracle (test infrastructure: checker only)."""
    from common import default_tables, pack
    from oracle_lib import Oracle
    tt, ct, _ = default_tables()
    orc = parity_sample.orc = getattr(parity_sample, "orc", None) or Oracle()
    sub = [seqs[i] for i in idx]
    b, o = pack(sub)
    cores = os.cpu_count() or 1
    if inner_threads > 1:
        want = orc.fold_batch(b, o, contra, False, tt, ct, gammas, inner_threads=inner_threads)
    else:
        want = orc.fold_batch(b, o, contra, False, tt, ct, gammas, n_threads=cores)
    ok = True
    for k, (logz, bpp, st) in enumerate(runner.results(idx)):
        lo, hi = int(want["bpp_offsets"][k]), int(want["bpp_offsets"][k + 1])
        ok = ok and np.float32(logz).view(np.uint32) == want["logz"][k].view(np.uint32)
        ok = ok and bool((bpp.view(np.uint32) == want["bpp"][lo:hi].view(np.uint32)).all())
        if st is not None:
            ok = ok and bool((st == want["structs"][:, o[k]:o[k + 1]]).all())
    return bool(ok)


def config_legs(h, dev, stream, args):
    """Short legs for every BASELINE.json config on one GPU, each with a bit-parity spot check (N = 1 only)."""
    import torch
    from common import default_tables, load_trnas, pack
    from oracle_lib import Oracle
    from rna_algos_b200 import _lib
    cores = os.cpu_count() or 1
    out = {}
    trnas = load_trnas()
    tiled = [trnas[i % 6] for i in range(args.nseq)]
    sweep = [float(np.float32(2.0) ** p) for p in range(-7, 11)]   # src/bin/centroid_fold.rs:9-11
    # configs[0]: Turner 2004 on the tRNA set; gamma = 2 and the reference's default sweep under CONTRAfold
    for key, contra, gam in (("c0_trna_turner_g1", False, [1.0]), ("c1_trna_contra_g2", True, [2.0]),
                             ("c1_trna_contra_sweep18", True, sweep)):
        r = FoldRunner(h, tiled, contra, gam, dev, stream)
        ms = r.timed(2, 1)
        out[key] = {"seq_s": round(r.n / ms * 1e3), "ms": round(ms, 1), "parity_ok": parity_sample(r, tiled, contra, gam, range(6))}
        del r
    # the same tRNA batch in the FAST_F32 numeric mode (exact log-space arithmetic, not bit-exact): rate and deviation
    def fast_leg(r):
        z0, p0 = r.t["logz"].clone(), r.t["bpp"].clone()
        h.set_numeric_mode("fast")
        try:
            msf = r.timed(2, 1)
        finally:
            h.set_numeric_mode("exact")
        present = p0 >= 0
        return {"seq_s": round(r.n / msf * 1e3), "ms": round(msf, 1),
                "max_dlogz": float(f"{(r.t['logz'] - z0).abs().max().item():.3g}"),
                "max_dbpp": float(f"{(r.t['bpp'] - p0)[present].abs().max().item():.3g}")}
    r = FoldRunner(h, tiled, True, [1.0], dev, stream)
    r.timed(1, 0)
    out["c1_trna_contra_g1_fast_f32"] = fast_leg(r)
    del r
    # configs[2]: the seeded Rfam-like families (50-500 nt), one GPU's share
    fam_seqs, fam_pairs = family_batch(args.rfam_nseq)
    r = FoldRunner(h, fam_seqs, True, [1.0], dev, stream)
    ms = r.timed(1, 1)
    lens = r.lens
    pick = [int(np.argmax(lens)), int(np.argmin(lens)), 0, len(fam_seqs) // 2, len(fam_seqs) - 1]
    out["c2_rfam_like_contra"] = {"seq_s": round(r.n / ms * 1e3), "ms": round(ms, 1), "n": r.n, "len": [int(lens.min()), int(lens.max())],
                                  "cells_s": round(cells_of(lens) / ms * 1e3), "parity_ok": parity_sample(r, fam_seqs, True, [1.0], pick)}
    out["c2_rfam_like_contra_fast_f32"] = fast_leg(r)
    del r
    # configs[3]: one random sequence at 1k / 2k / 4k nt, both models; parity at 1024 here, 2048 / 4096 in tests/
    longs = {}
    for L, seed in ((1024, 1), (2048, 2), (4096, 3)):
        seq = np.random.default_rng(seed).integers(0, 4, size=L).astype(np.uint8)
        ent = {}
        for contra in (False, True):
            r = FoldRunner(h, [seq], contra, [1.0], dev, stream)
            ms1 = r.timed(1, 1)
            if L < 4096:
                ms1 = min(ms1, r.timed(1, 0))   # (single calls: best of two, the 4096-nt one takes seconds)
            ent["contra_ms" if contra else "turner_ms"] = round(ms1, 1)
            if L == 1024:
                ent["parity_ok"] = ent.get("parity_ok", True) and parity_sample(r, [seq], contra, [1.0], [0], inner_threads=cores)
            if contra:   # FAST numeric mode on the same sequence (f32 warp-shuffle reductions), with its deviation
                exact_logz = float(r.t["logz"][0])
                h.set_numeric_mode("fast")
                try:
                    ent["fast_ms"] = round(min(r.timed(1, 1), r.timed(1, 0), r.timed(1, 0)), 1)   # (best of three single calls)
                    ent["fast_dlogz"] = float(f"{abs(float(r.t['logz'][0]) - exact_logz):.3g}")
                finally:
                    h.set_numeric_mode("exact")
            del r
        longs[str(L)] = ent
    out["c3_long"] = longs
    # configs[4]: Durbin forward-backward on the intra-family pairs: device-resident (CUDA events) and through host buffers
    b, o = pack(fam_seqs)
    _, _, at = default_tables()
    lens64 = np.diff(o.astype(np.int64))
    sizes = (lens64[fam_pairs[:, 0]] + 2) * (lens64[fam_pairs[:, 1]] + 2)
    po = np.zeros(len(fam_pairs) + 1, dtype=np.uint64)
    po[1:] = np.cumsum(sizes)
    d = {"bases": torch.from_numpy(b).to(dev), "off": torch.from_numpy(o.view(np.int32)).to(dev),
         "pairs": torch.from_numpy(fam_pairs.view(np.int32).reshape(-1)).to(dev), "po": torch.from_numpy(po.view(np.int64)).to(dev),
         "out": torch.empty(int(po[-1]), dtype=torch.float32, device=dev)}
    db = _lib.DurbinBatchDev()
    db.h_offsets = o.ctypes.data; db.h_pairs = fam_pairs.ctypes.data
    db.d_bases = d["bases"].data_ptr(); db.d_offsets = d["off"].data_ptr(); db.d_pairs = d["pairs"].data_ptr()
    db.d_prob_offsets = d["po"].data_ptr(); db.n_seqs = len(fam_seqs); db.n_pairs = len(fam_pairs); db.max_len = int(lens64.max())
    db.d_out_probs = d["out"].data_ptr()
    sptr = C.c_void_p(stream.cuda_stream)

    def dstep():
        rc = h.lib.rna_durbin_batch_dev(h.h, C.byref(db), sptr)
        if rc:
            raise RuntimeError(f"rna_durbin_batch_dev rc={rc}: {h.lib.rna_last_error(h.h)}")
    dstep()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    dstep()
    ev1.record(stream)
    torch.cuda.synchronize()
    dms = ev0.elapsed_time(ev1)
    # end to end from host buffers; the match-probability matrices (GBs) land in a page-locked buffer, which the kernels
    # write directly over PCIe (a pageable buffer takes the staged path and is several times slower)
    pinned_out = torch.empty(int(po[-1]), dtype=torch.float32, pin_memory=True).numpy()
    res = h.durbin_batch(b, o, fam_pairs, out=pinned_out)
    t1 = time.perf_counter()
    res = h.durbin_batch(b, o, fam_pairs, out=pinned_out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t1
    k = min(len(fam_pairs), 48)
    want = Oracle().durbin_batch(b, o, fam_pairs[:k], at, n_threads=cores)
    hi = int(res["prob_offsets"][k])
    dev_ok = bool((d["out"][:hi].cpu().numpy().view(np.uint32) == want["probs"].view(np.uint32)).all())
    out["c4_durbin_family_pairs"] = {"pairs_s": round(len(fam_pairs) / dms * 1e3), "ms": round(dms, 1), "n": int(len(fam_pairs)),
                                     "cells_s": round(int(sizes.sum()) / dms * 1e3), "e2e_pairs_s": round(len(fam_pairs) / dt),
                                     "e2e_out_bytes": int(po[-1]) * 4,
                                     "parity_ok": dev_ok and bool((res["probs"][:hi].view(np.uint32) == want["probs"].view(np.uint32)).all())}
    return out


def strong_leg(h, dev, stream, args, world, rank, dist):
    """Strong scaling on a ragged batch: the fixed seeded Rfam-like global batch, LPT-partitioned over the ranks."""
    import torch
    from rna_algos_b200.api import fold_cost, partition_lpt
    seqs_all, _ = family_batch(args.rfam_nseq)
    part = partition_lpt(fold_cost([len(s) for s in seqs_all]), world)
    mine = [s for s, p in zip(seqs_all, part) if p == rank]
    r = FoldRunner(h, mine, True, [1.0], dev, stream)
    r.timed(1, 0)
    if world > 1:
        dist.barrier()
    ms = r.timed(1, 0)
    busy = torch.tensor([ms], dtype=torch.float64, device=dev)
    allb = [torch.zeros_like(busy) for _ in range(world)]
    if world > 1:
        dist.all_gather(allb, busy)
    else:
        allb = [busy]
    busy_ms = [float(x.item()) for x in allb]
    return {"workload": f"{len(seqs_all)} seeded Rfam-like sequences (50-500 nt), fixed global batch, LPT on L^3+500L^2",
            "seq_s": round(len(seqs_all) / max(busy_ms) * 1e3), "busy_ms": [round(x, 1) for x in busy_ms],
            "imbalance_max_over_mean": round(max(busy_ms) / (sum(busy_ms) / len(busy_ms)), 3)}


# ----------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    from common import default_tables, pack
    from rna_algos_b200 import _lib
    from rna_algos_b200.api import Handle, bpp_offsets_of, fold_cost, partition_lpt

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gammas = [float(g) for g in args.gammas.split(",")]
    # global batch = world x nseq, length-balanced LPT partition over ranks (no collective on the data path)
    seqs_all, contra, desc = make_workload(args.workload, args.nseq * world, length=args.length)
    desc = desc.replace(f"{args.nseq * world} seqs/GPU", f"{args.nseq} seqs/GPU")
    if world > 1:
        part = partition_lpt(fold_cost([len(s) for s in seqs_all]), world)
        seqs = [s for s, p in zip(seqs_all, part) if p == rank]
    else:
        seqs = seqs_all
    bases, offsets = pack(seqs)
    n = len(seqs)
    lens = np.diff(offsets.astype(np.int64))
    total = int(offsets[-1])
    bpp_off = bpp_offsets_of(offsets)
    bpp_total = int(bpp_off[-1])
    ng = len(gammas)

    tt, ct, at = default_tables()
    h = Handle(local, tt, ct, at)
    lib = h.lib

    # ---- device-resident buffers (torch = device memory + streams only) -------------------------------
    d_bases = torch.from_numpy(bases).to(dev)
    d_offsets = torch.from_numpy(offsets.view(np.int32)).to(dev)
    d_bppoff = torch.from_numpy(bpp_off.view(np.int64)).to(dev)
    d_gammas = torch.tensor(gammas, dtype=torch.float32, device=dev)
    d_logz = torch.empty(n, dtype=torch.float32, device=dev)
    d_bpp = torch.empty(max(bpp_total, 1), dtype=torch.float32, device=dev)
    d_structs = torch.empty(max(ng * total, 1), dtype=torch.uint8, device=dev)
    d_ea = torch.empty(max(ng * n, 1), dtype=torch.float32, device=dev)
    fb = _lib.FoldBatchDev()
    fb.h_offsets = offsets.ctypes.data
    fb.d_bases = d_bases.data_ptr(); fb.d_offsets = d_offsets.data_ptr(); fb.d_bpp_offsets = d_bppoff.data_ptr()
    fb.n_seqs = n; fb.total_len = total; fb.max_len = int(lens.max())
    fb.model = _lib.MODEL_CONTRA if contra else _lib.MODEL_TURNER
    fb.allows_short_hairpins = 0
    fb.d_gammas = d_gammas.data_ptr(); fb.n_gammas = ng
    fb.d_out_logz = d_logz.data_ptr(); fb.d_out_bpp = d_bpp.data_ptr()
    fb.d_out_structs = d_structs.data_ptr(); fb.d_out_expect_acc = d_ea.data_ptr()
    # a dedicated non-default stream: the library treats a NULL stream as "use the handle's own stream", and
    # CUDA events only see the stream they are recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)
    assert sptr.value, "expected a non-default stream"

    def step_dev():
        rc = lib.rna_mccaskill_centroid_batch_dev(h.h, C.byref(fb), sptr)
        if rc:
            raise RuntimeError(f"rna_mccaskill_centroid_batch_dev rc={rc}: {lib.rna_last_error(h.h)}")

    for _ in range(max(args.warmup, 3)):
        step_dev()
    barrier()
    launches0 = h.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        step_dev()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = h.stats()["kernel_launches"] - launches0
    # quick self-check of the timed result (not timed): logZ finite, every structure byte is . ( or )
    assert torch.isfinite(d_logz).all(), "non-finite logZ"
    sb = d_structs[: ng * total]
    assert bool(((sb == 46) | (sb == 40) | (sb == 41)).all()), "structure bytes corrupted"

    # ---- end to end through the host-buffer C-ABI call: pinned host in / out ----------------------------
    def pinned(shape, dtype):
        return torch.empty(shape, dtype=dtype).pin_memory().numpy()
    pb = pinned(bases.shape[0], torch.uint8); pb[:] = bases
    out = {"logz": pinned(n, torch.float32), "bpp": pinned(max(bpp_total, 1), torch.float32)[:bpp_total],
           "structs": pinned((ng, total), torch.uint8), "expect_acc": pinned((ng, n), torch.float32)}
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        h.fold_batch(pb, offsets, contra, False, gammas, out=out)
    st = h.stats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h.fold_batch(pb, offsets, contra, False, gammas, out=out)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    barrier()

    # ---- reduce over ranks: max time, sum of units ---------------------------------------------------------
    red = torch.tensor([ms, e2e_dt], dtype=torch.float64, device=dev)
    cnt = torch.tensor([n, cells_of(lens), algorithmic_bytes(lens, ng), launches, st["h2d_bytes"], st["d2h_bytes"]],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(red, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, e2e_dt = (float(x) for x in red.tolist())
    n_all, cells_all, abytes_all, launches_all, h2d_all, d2h_all = (float(x) for x in cnt.tolist())

    strong = None
    if world > 1 and not args.no_configs:
        strong = strong_leg(h, dev, stream, args, world, rank, dist)
    own_peaks = measure_peaks(local) if rank == 0 else None
    configs = None
    if world == 1 and not args.no_configs:
        configs = config_legs(h, dev, stream, args)

    if rank == 0:
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, peak_src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            hbm_peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        sec_per_step = ms / 1e3 / args.steps
        value = n_all * args.steps / (ms / 1e3)
        # cpu baseline (rank 0, N = 1 only): oracle on all host cores over a bounded sample
        cpu = None
        lse_per_seq = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            crate, cn, cdt, lse_per_seq = cpu_sample(seqs, contra, gammas, cores, args.cpu_seconds)
            cpu = {"value": crate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {cn} sequences of the same workload, {cdt:.1f} s wall (oracle/oracle.c port of the "
                             f"reference algorithm, one sequence per task on {cores} threads)"}
        per_gpu_bytes = abytes_all / world
        hbm_ach = per_gpu_bytes / sec_per_step / 1e9
        # DRAM bytes per step of the same command, from the committed ncu capture (profiles/), if it matches
        traffic = None
        for tname in ("r2_traffic.json", "r1_traffic.json"):
            try:
                tj = json.load(open(os.path.join(ROOT, "profiles", tname)))
                if tj.get("workload") == args.workload and tj.get("nseq_per_gpu") == args.nseq and tj.get("gammas") == gammas:
                    traffic = tj["dram_bytes_per_step"]
                    break
            except Exception:
                pass
        # algorithmic work: logsumexp terms per sequence, counted by the instrumented oracle (exact for the tiled tRNA
        # workload: its 6 distinct sequences; else from the CPU sample above)
        if lse_per_seq is None and args.workload in ("trna_contra", "trna_turner"):
            from common import load_trnas
            _, r6 = cpu_run(load_trnas(), contra, gammas, min(6, os.cpu_count() or 1))
            lse_per_seq = r6["lse_terms"] / 6
        roof = {"bound": "fp32_issue", "achieved": None, "peak": None, "unit": "G FP32 thread-instr/s per GPU", "frac": None,
                "traffic": traffic, "kernel": f"fold_kernel2<{'CONTRA' if contra else 'TURNER'}, SMEM>"}
        if lse_per_seq is not None:
            ach = lse_per_seq * value / world * FP32_INSTR_PER_LSE / 1e9
            if own_peaks:
                peak, src = own_peaks["fp32_nofma_ginstr_s"], "measured in-run (tools/peak_microbench.cu: separate FMUL/FADD, no FMA)"
            else:
                sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
                peak, src = 148 * 128 * sm_mhz * 1e-3, "nominal 148 SM x 128 lanes x clock (libpeaks.so missing)"
            roof.update({"achieved": ach, "peak": peak, "frac": ach / peak, "peak_source": src,
                         "lse_terms_per_seq": lse_per_seq, "fp32_instr_per_lse": FP32_INSTR_PER_LSE,
                         "note": "achieved = counted logsumexp terms/s x 14 FP32 instructions (SURVEY 8(d)); the reference-exact "
                                 "chains are latency-bound, see DESIGN.md 3.3"})
        roof["hbm"] = {"achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "frac": hbm_ach / hbm_peak, "peak_source": peak_src,
                       "algorithmic_bytes_per_step": per_gpu_bytes}
        if own_peaks:
            roof["measured_peaks"] = {k: round(v, 1) for k, v in own_peaks.items()}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (bundled tRNAs tiled)",
            "config": {"workload": desc, "gammas": gammas, "nseq_per_gpu": args.nseq},
            "notes": {"numeric_mode": "reference-exact f32",
                      "l2": f"per-step output working set {4 * bpp_total / 1e6:.0f} MB/GPU vs 126 MB L2 "
                            f"({'exceeds L2, no flush needed' if 4 * bpp_total > 126e6 else 'smaller than L2'})",
                      "partition": "LPT on L^3+500L^2 over ranks, no collective"},
            "cells_per_s": cells_all * args.steps / (ms / 1e3),
            "e2e": {"value": n_all * e2e_steps / e2e_dt, "unit": UNIT, "h2d_bytes_per_step": h2d_all,
                    "d2h_bytes_per_step": d2h_all, "steps": e2e_steps,
                    "path": "rna_mccaskill_centroid_batch (host buffers, pinned)"},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        if strong is not None:
            line["strong"] = strong
        if configs is not None:
            line["configs"] = configs   # (last key: the driver keeps the tail of the line)
        print(json.dumps(line), flush=True)
    h.close()
    if world > 1:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------------
# Durbin pair-HMM arm (BASELINE configs[4]; secondary metric, same JSON shape): --workload durbin
# ----------------------------------------------------------------------------------------------------
def durbin_arm(args):
    """All 15 pairs of the 6 bundled tRNAs (the reference's own Durbin benchmark set, benches/benches.rs:58-93)
    tiled to --nseq pairs per GPU, CONTRAlign v2.01 scores.  value = pairs/s device-resident; e2e = host buffers."""
    import torch
    from common import default_tables, load_trnas, pack
    from oracle_lib import Oracle
    from rna_algos_b200 import _lib
    from rna_algos_b200.api import Handle
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device")
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    tt, ct, at = default_tables()
    h = Handle(0, None, None, at)
    lib = h.lib
    seqs = load_trnas()
    bases, offsets = pack(seqs)
    base_pairs = [(a, b) for a in range(6) for b in range(a + 1, 6)]
    npairs = args.nseq
    pairs = np.array([base_pairs[i % 15] for i in range(npairs)], dtype=np.uint32)
    lens = np.diff(offsets.astype(np.int64))
    sizes = (lens[pairs[:, 0]] + 2) * (lens[pairs[:, 1]] + 2)
    po = np.zeros(npairs + 1, dtype=np.uint64)
    po[1:] = np.cumsum(sizes)
    d_bases = torch.from_numpy(bases).to(dev)
    d_off = torch.from_numpy(offsets.view(np.int32)).to(dev)
    d_pairs = torch.from_numpy(pairs.view(np.int32).reshape(-1)).to(dev)
    d_po = torch.from_numpy(po.view(np.int64)).to(dev)
    d_out = torch.empty(int(po[-1]), dtype=torch.float32, device=dev)
    db = _lib.DurbinBatchDev()
    db.h_offsets = offsets.ctypes.data; db.h_pairs = pairs.ctypes.data
    db.d_bases = d_bases.data_ptr(); db.d_offsets = d_off.data_ptr(); db.d_pairs = d_pairs.data_ptr()
    db.d_prob_offsets = d_po.data_ptr(); db.n_seqs = 6; db.n_pairs = npairs; db.max_len = int(lens.max())
    db.d_out_probs = d_out.data_ptr()
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    sptr = C.c_void_p(stream.cuda_stream)

    def step():
        rc = lib.rna_durbin_batch_dev(h.h, C.byref(db), sptr)
        if rc:
            raise RuntimeError(f"rna_durbin_batch_dev rc={rc}: {lib.rna_last_error(h.h)}")
    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop()
    host_out = torch.empty(int(po[-1]), dtype=torch.float32).pin_memory().numpy()
    for _ in range(2):
        h.durbin_batch(bases, offsets, pairs, out=host_out)
    st = h.stats()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(e2e_steps):
        h.durbin_batch(bases, offsets, pairs, out=host_out)
    e2e_dt = time.perf_counter() - t0
    # parity spot check + cpu baseline on a bounded sample (oracle port, all cores)
    cores = os.cpu_count() or 1
    orc = Oracle()
    ncpu = min(npairs, 15 * 64)
    t0 = time.perf_counter()
    want = orc.durbin_batch(bases, offsets, pairs[:ncpu], at, n_threads=cores)
    cdt = time.perf_counter() - t0
    assert (host_out[: int(po[ncpu])].view(np.uint32) == want["probs"].view(np.uint32)).all(), "Durbin parity"
    cells = int(sizes.sum())
    abytes = 4 * cells + 8 * npairs + 8 * npairs
    sec = ms / 1e3 / args.steps
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        hbm_peak = 6650.0
    line = {
        "metric": "durbin_match_prob_pairs_per_s", "value": npairs / sec, "unit": "pairs/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * sec, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (bundled tRNA pairs tiled)",
        "config": {"workload": f"15 tRNA pairs of assets/sampled_trnas.fa tiled to {npairs} pairs, CONTRAlign v2.01, sentinel-padded"},
        "cells_per_s": cells / sec,
        "e2e": {"value": npairs * e2e_steps / e2e_dt, "unit": "pairs/s", "h2d_bytes_per_step": st["h2d_bytes"],
                "d2h_bytes_per_step": st["d2h_bytes"], "steps": e2e_steps, "path": "rna_durbin_batch (host buffers)"},
        "gpu_launches": args.steps, "clocks": clocks,
        "roofline": {"bound": "hbm", "achieved": abytes / sec / 1e9, "peak": hbm_peak, "unit": "GB/s",
                     "frac": abytes / sec / 1e9 / hbm_peak, "traffic": None, "kernel": "durbin_kernel",
                     "note": "algorithmic bytes = match-probability matrices out (4 B per cell) + pair list"},
        "cpu_baseline": {"value": ncpu / cdt, "unit": "pairs/s", "cores": cores, "kind": "port",
                         "sample": f"first {ncpu} pairs, {cdt:.2f} s wall (oracle port, bit-identical to the GPU result)"},
    }
    print(json.dumps(line), flush=True)
    h.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="trna_contra")
    ap.add_argument("--nseq", type=int, default=24576, help="sequences per GPU per step")
    ap.add_argument("--length", type=int, default=76)
    ap.add_argument("--gammas", default="1.0")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--ref-step-seconds", type=float, default=3.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config legs (N=1) / the strong-scaling leg (N>1)")
    ap.add_argument("--rfam-nseq", type=int, default=4096, help="sequences of the seeded Rfam-like batch (configs[2], [4], strong)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload == "durbin":
        durbin_arm(args)
    else:
        gpu_arm(args)


if __name__ == "__main__":
    main()
