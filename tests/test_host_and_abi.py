"""CPU-side checks: the C-ABI library exports every symbol the header declares and binds with the declared blob
sizes, the pure-host entry points (validation, LPT partition, table helpers), the Python mirror of the reference's
input rules, the oracle against the committed golden vectors and the reference's own range assertion, and the
multi-rank sharding logic on two gloo ranks.  No GPU, no compute calls into the CUDA path."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, default_tables, load_trnas, pack, random_seqs
from oracle_lib import Oracle
from rna_algos_b200 import _lib
from rna_algos_b200 import tables as T

HEADER = os.path.join(ROOT, "include", "rna_algos_b200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(rna_[a-z0-9_]+)\s*\(", src))
    return sorted(n for n in names if n not in ("rna_bpp_len", "rna_bpp_index", "rna_sums_len", "rna_sums_index"))   # static inline helpers


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = header_functions()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/rna_algos_b200.h but not exported"
    assert set(_lib.exported_names()) <= set(declared)
    assert lib.rna_version().startswith(b"rna_algos_b200")
    assert lib.rna_sizeof_turner_tables() == C.sizeof(T.TurnerTables)
    assert lib.rna_sizeof_contra_tables() == C.sizeof(T.ContraTables)
    assert lib.rna_sizeof_align_tables() == C.sizeof(T.AlignTables)


def test_rust_shim_binds_only_exported_symbols_with_matching_blob_sizes():
    """rust/rna_algos_b200_shim cannot be compiled here (no Rust toolchain): at least every extern "C" name it declares
    must be exported by the library, and the #[repr(C)] table structs must add up to the C sizes."""
    lib = _lib.load()
    src = open(os.path.join(ROOT, "rust", "rna_algos_b200_shim", "src", "ffi.rs")).read()
    names = re.findall(r"pub fn (rna_[a-z0-9_]+)\s*\(", src)
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), name

    def struct_bytes(name):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % name, src, re.S).group(1)
        alias = {"T4": 256 * 4, "T3": 64 * 4, "f32": 4, "i32": 4, "u8": 1, "RnaSpecialHairpin": 20}

        def size(t):
            t = t.strip()
            m = re.match(r"\[(.*);\s*([A-Z_0-9a-z]+)\]$", t)
            if m:
                n = {"RNA_LOOP_TABLE_LEN": 31, "RNA_MAX_SPECIAL_HAIRPINS": 128, "RNA_MAX_SPECIAL_HAIRPIN_LEN": 12}.get(m.group(2))
                return size(m.group(1)) * (n if n is not None else int(m.group(2)))
            return alias[t]
        return sum(size(f.split(":", 1)[1].strip().rstrip(",")) for f in body.split(",\n") if ":" in f)

    assert struct_bytes("RnaTurnerTables") == lib.rna_sizeof_turner_tables()
    assert struct_bytes("RnaContraTables") == lib.rna_sizeof_contra_tables()
    assert struct_bytes("RnaAlignTables") == lib.rna_sizeof_align_tables()


def test_no_cpu_fallback_without_device():
    """Without a CUDA device the product refuses to compute (RNA_ERR_NO_DEVICE), it never falls back."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.rna_create(0, C.byref(h)) == 8
    from rna_algos_b200.api import Handle, RnaError
    with pytest.raises(RnaError):
        Handle(0)


def test_validate_bases_codes():
    lib = _lib.load()
    b = np.array([0, 1, 2, 3, 3, 2], dtype=np.uint8)
    o = np.array([0, 4, 6], dtype=np.uint32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.rna_validate_bases(p(b), p(o), 2) == 0
    b2 = b.copy(); b2[2] = 4
    assert lib.rna_validate_bases(p(b2), p(o), 2) == 2            # RNA_ERR_INVALID_BASE (reference: bytes2seq panics)
    assert lib.rna_validate_bases(p(b), p(np.array([0, 0, 6], dtype=np.uint32)), 2) == 3   # RNA_ERR_EMPTY_SEQ
    assert lib.rna_validate_bases(p(b), p(np.array([0, 5, 4], dtype=np.uint32)), 2) == 1   # RNA_ERR_BAD_ARG


def test_partition_lpt_balanced_and_deterministic():
    from rna_algos_b200.api import fold_cost, partition_lpt
    rng = np.random.default_rng(1)
    lens = rng.integers(50, 500, size=3000)
    costs = fold_cost(lens)
    for parts in (1, 2, 4, 8):
        part = partition_lpt(costs, parts)
        assert part.min() == 0 and part.max() == parts - 1
        loads = np.array([costs[part == k].sum() for k in range(parts)], dtype=np.float64)
        assert loads.max() / loads.mean() < 1.01                   # LPT on 3000 units is near-perfect
        assert (partition_lpt(costs, parts) == part).all()


def test_bytes2seq_rules():
    """src/utils.rs:562-577: ACGU either case; anything else (incl. T, N) is an error."""
    from rna_algos_b200.api import RnaError, bytes2seq, get_fold_str
    assert bytes2seq("ACGUacgu").tolist() == [0, 1, 2, 3, 0, 1, 2, 3]
    for bad in ("ACGT", "ACGN", "AC-GU", "AC GU"):
        with pytest.raises(RnaError):
            bytes2seq(bad)
    assert get_fold_str([(0, 5), (1, 4)], 6) == "((..))"


def test_contra_accumulate_matches_oracle():
    lib = _lib.load()
    orc = Oracle()
    a = T.random_contra_tables(7)
    b = T.random_contra_tables(7)
    lib.rna_contra_tables_accumulate(C.byref(a))
    orc.lib.orc_contra_accumulate(C.byref(b))
    assert bytes(a) == bytes(b)


def test_numerics_error_bounds():
    """The polynomial kernels reproduce the figures measured on the reference."""
    orc = Oracle()
    x = np.linspace(0, 11.862479, 20001, dtype=np.float32)
    err = max(abs(orc.lib.orc_ln_exp_1p(float(v)) - np.log1p(np.exp(np.float64(v)))) for v in x[::10])
    assert 5e-6 < err < 9e-6
    assert orc.lib.orc_expf(-9.92) == 0.0 and orc.lib.orc_expf(-9.90) > 0.0
    assert orc.lib.orc_logsumexp(float("-inf"), 1.5) == 1.5 and orc.lib.orc_logsumexp(2.5, float("-inf")) == 2.5


@pytest.mark.parametrize("contra", [False, True])
def test_oracle_matches_golden_vectors(contra):
    g = np.load(os.path.join(ROOT, "tests", "golden", "trna_oracle.npz"))
    tt, ct, at = default_tables()
    bases, offsets = pack(load_trnas())
    r = Oracle().fold_batch(bases, offsets, contra, False, tt, ct, g["gammas"].tolist(), n_threads=4)
    k = "contra" if contra else "turner"
    assert (r["logz"].view(np.uint32) == g[k + "_logz"].view(np.uint32)).all()
    assert (r["bpp"].view(np.uint32) == g[k + "_bpp"].view(np.uint32)).all()
    assert (r["structs"] == g[k + "_structs"]).all()
    assert (r["expect_acc"].view(np.uint32) == g[k + "_expect_acc"].view(np.uint32)).all()
    # the reference's only assertion (tests/tests.rs:33,38): every probability in [-0.001, 1.001)
    p = r["bpp"][r["bpp"] != T.BPP_ABSENT]
    assert (p >= -0.001).all() and (p < 1.001).all()
    # structures are balanced dot-brackets of the right length
    for row in r["structs"]:
        for s in range(6):
            seg = row[offsets[s]:offsets[s + 1]].tobytes().decode()
            assert seg.count("(") == seg.count(")") and set(seg) <= set(".()")


def test_oracle_durbin_golden_and_range():
    g = np.load(os.path.join(ROOT, "tests", "golden", "trna_oracle.npz"))
    _, _, at = default_tables()
    bases, offsets = pack(load_trnas())
    pairs = np.array([(a, b) for a in range(6) for b in range(a + 1, 6)], dtype=np.uint32)
    d = Oracle().durbin_batch(bases, offsets, pairs, at, n_threads=4)
    assert (d["probs"].view(np.uint32) == g["durbin_probs"].view(np.uint32)).all()
    assert (d["probs"] >= -0.001).all() and (d["probs"] < 1.001).all()       # tests/tests.rs:74


_RANK_SCRIPT = r"""
import os, sys, json
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], 'tests'))
from rna_algos_b200.api import fold_cost, partition_lpt
import bench
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
seqs, contra, desc = bench.make_workload('rfam_synth_contra', 200 * world)
part = partition_lpt(fold_cost([len(s) for s in seqs]), world)
mine = [i for i, p in enumerate(part) if p == rank]
# every unit is owned by exactly one rank; loads are balanced; no data-path collective is needed
owned = torch.zeros(len(seqs), dtype=torch.int32); owned[mine] = 1
dist.all_reduce(owned)
load = torch.tensor([float(sum(int(fold_cost(len(seqs[i]))) for i in mine))], dtype=torch.float64)
loads = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
dist.all_gather(loads, load)
t = torch.tensor([1.0 + rank], dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({'covered': bool((owned == 1).all()), 'imbalance': max(l.item() for l in loads) / (sum(l.item() for l in loads) / world),
                      'max_reduce': t.item(), 'cells': bench.cells_of(np.array([len(s) for s in seqs]))}))
dist.destroy_process_group()
"""


def test_two_rank_sharding_gloo(tmp_path):
    """The N>1 path of bench.py on CPU: LPT shards over 2 gloo ranks, max-over-ranks reduction."""
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT],
                       capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    out = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert out["covered"] and out["imbalance"] < 1.05 and out["max_reduce"] == 2.0 and out["cells"] > 0


def test_reference_arm_runs_on_cpu():
    """bench.py --impl reference (the CPU arm) prints one JSON line with the contract's keys."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--ref-step-seconds", "0.5"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "cpu_baseline", "e2e", "config"):
        assert k in line
    assert line["impl"] == "reference" and line["value"] > 0 and line["cpu_baseline"]["kind"] == "port"
