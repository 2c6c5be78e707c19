"""Command-line front ends (cli/rna_cli.cpp): the reference's three programs, options and text formats
(src/bin/mccaskill_algo.rs, src/bin/centroid_fold.rs, src/bin/durbin_algo.rs) over the C ABI.
CPU tests cover parsing, formatting and error behaviour; GPU tests diff whole output files against text built
from the oracle's results."""
import os
import subprocess

import numpy as np
import pytest

from common import ROOT, TRNA_FASTA, default_tables, load_trnas, pack

CLI = os.path.join(ROOT, "bin", "rna_algos_b200")


@pytest.fixture(scope="module")
def cli():
    subprocess.run(["make", "-C", ROOT, "cli"], check=True, capture_output=True)
    assert os.path.exists(CLI)
    return CLI


def run(cli, *args):
    return subprocess.run([cli, *args], capture_output=True, text=True)


def rust_f32(x) -> str:
    """Rust's `{}` for f32: shortest round-trip decimal, positional."""
    return np.format_float_positional(np.float32(x), unique=True, trim="-")


def test_help_and_usage(cli):
    for prog in ("mccaskill_algo", "centroid_fold", "durbin_algo"):
        r = run(cli, prog, "-h")
        assert r.returncode == 0 and "--input_file_path" in r.stdout
    assert "--centroid_threshold" in run(cli, "centroid_fold", "-h").stdout
    assert "--uses_contra_model" not in run(cli, "durbin_algo", "-h").stdout
    r = run(cli, "mccaskill_algo", "-i", "x.fa")
    assert r.returncode == 1 and "-i and -o are required" in r.stderr


def test_float_format_is_rusts_display(cli):
    rng = np.random.default_rng(5)
    vals = np.concatenate([
        np.array([0.0, 1.0, 0.5, 0.0078125, 1024.0, 1e-7, 0.1, 0.3, 0.99999994, 5e-5, 1.4e-45, 3.4028235e38], dtype=np.float32),
        rng.uniform(0, 1, 200).astype(np.float32), np.exp(rng.uniform(-30, 0, 200)).astype(np.float32)])
    r = run(cli, "_fmt", *[f"{b:08x}" for b in vals.view(np.uint32)])
    got = r.stdout.split()
    assert got == [rust_f32(v) for v in vals]
    assert all(np.float32(g) == v for g, v in zip(got, vals))   # round trip


def test_bad_base_is_an_error_not_a_crash(cli, tmp_path):
    fa = tmp_path / "bad.fa"
    fa.write_text(">x\nACGUN\n")
    r = run(cli, "mccaskill_algo", "-i", str(fa), "-o", str(tmp_path / "o.dat"))
    assert r.returncode == 1 and "not one of ACGU" in r.stderr
    r = run(cli, "durbin_algo", "-i", str(tmp_path / "missing.fa"), "-o", str(tmp_path / "o.dat"))
    assert r.returncode == 1 and "cannot open" in r.stderr


def test_fasta_reader_semantics(cli, tmp_path):
    """bio::io::fasta as the reference uses it: id = first word of the header, the sequence is every following line up
    to the next header with line ends removed, a/c/g/u in either case (src/utils.rs:562-577)."""
    fa = tmp_path / "in.fa"
    fa.write_bytes(b">first some description\nACGU\nacgu\r\n\nGG\n>second\tx\nu\n>third\nAcGuAcGu")
    r = run(cli, "_parse", str(fa))
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == ["first 10 ACGUACGUGG", "second 1 U", "third 8 ACGUACGU"]
    empty = tmp_path / "empty.fa"
    empty.write_text(">only_header\n")
    r = run(cli, "_parse", str(empty))
    assert r.returncode == 1 and "has no sequence" in r.stderr
    nohdr = tmp_path / "nohdr.fa"
    nohdr.write_text("ACGU\n")
    r = run(cli, "_parse", str(nohdr))
    assert r.returncode == 1 and "before the first" in r.stderr
    t = tmp_path / "t.fa"
    t.write_text(">dna\nACGT\n")
    r = run(cli, "_parse", str(t))
    assert r.returncode == 1 and "'T'" in r.stderr      # the reference panics on T as well


def test_no_device_fails_loudly(cli, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    r = run(cli, "centroid_fold", "-i", TRNA_FASTA, "-o", str(tmp_path / "out"), "--standin-tables")
    assert r.returncode == 1 and "no CPU path" in r.stderr


def test_standin_tables_need_an_explicit_opt_in(cli, tmp_path):
    """turner2004.tbl / contrafold_v202.tbl are names reserved for the genuine rna-ss-params values: without --tables
    (or $RNA_ALGOS_B200_TABLES) and without --standin-tables the fold programs refuse to run."""
    env = {k: v for k, v in os.environ.items() if k != "RNA_ALGOS_B200_TABLES"}
    r = subprocess.run([cli, "mccaskill_algo", "-i", TRNA_FASTA, "-o", str(tmp_path / "o.dat")], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "no score tables" in r.stderr and "--standin-tables" in r.stderr


# ---- GPU: whole-file parity -----------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("contra", [False, True])
def test_mccaskill_algo_output(cli, tmp_path, contra):
    from oracle_lib import Oracle
    tt, ct, _ = default_tables()
    out = tmp_path / "bpp.dat"
    r = run(cli, "mccaskill_algo", "-i", TRNA_FASTA, "-o", str(out), *(["-c"] if contra else []), "-t", "3", "--standin-tables")
    assert r.returncode == 0, r.stderr
    seqs = load_trnas()
    want = ("# Format = >{RNA sequence id} {line break} {basepairing left nucleotide}, {basepairing right nucleotide}, "
            "{basepairing probability} ...")
    orc = Oracle()
    for s, seq in enumerate(seqs):
        bpp, _ = orc.mccaskill(seq, contra, False, tt, ct)
        L = len(seq)
        want += f"\n\n>{s}\n"
        k = 0
        for i in range(L - 1):
            for j in range(i + 1, L):
                if bpp[k] != -1.0:
                    want += f"{i},{j},{rust_f32(bpp[k])} "
                k += 1
    assert out.read_text() == want


@pytest.mark.gpu
def test_centroid_fold_output(cli, tmp_path):
    from oracle_lib import Oracle
    tt, ct, _ = default_tables()
    seqs = load_trnas()
    bases, offsets = pack(seqs)
    orc = Oracle()
    # the reference's sweep 2^-7 .. 2^10 (no -g), CONTRAfold
    outdir = tmp_path / "sweep"
    r = run(cli, "centroid_fold", "-i", TRNA_FASTA, "-o", str(outdir), "-c", "--standin-tables")
    assert r.returncode == 0, r.stderr
    gammas = [float(np.float32(2.0) ** p) for p in range(-7, 11)]
    want = orc.fold_batch(bases, offsets, True, False, tt, ct, gammas, n_threads=4)
    names = sorted(os.listdir(outdir))
    assert names == sorted(f"centroid_threshold={rust_f32(g)}.fa" for g in gammas)
    for g, gamma in enumerate(gammas):
        text = "\n".join(f">{s}\n" + bytes(want["structs"][g, offsets[s]:offsets[s + 1]]).decode() for s in range(len(seqs)))
        assert (outdir / f"centroid_threshold={rust_f32(gamma)}.fa").read_text() == text
    # one threshold, Turner
    outdir = tmp_path / "one"
    r = run(cli, "centroid_fold", "-i", TRNA_FASTA, "-o", str(outdir), "-g", "2", "--standin-tables")
    assert r.returncode == 0, r.stderr
    assert os.listdir(outdir) == ["centroid_threshold=2.fa"]
    want = orc.fold_batch(bases, offsets, False, False, tt, ct, [2.0], n_threads=4)
    text = "\n".join(f">{s}\n" + bytes(want["structs"][0, offsets[s]:offsets[s + 1]]).decode() for s in range(len(seqs)))
    assert (outdir / "centroid_threshold=2.fa").read_text() == text


@pytest.mark.gpu
def test_durbin_algo_output(cli, tmp_path):
    from oracle_lib import Oracle
    _, _, at = default_tables()
    seqs = load_trnas()
    bases, offsets = pack(seqs)
    out = tmp_path / "match.dat"
    r = run(cli, "durbin_algo", "-i", TRNA_FASTA, "-o", str(out))
    assert r.returncode == 0, r.stderr
    pairs = np.array([(a, b) for a in range(len(seqs)) for b in range(a + 1, len(seqs))], dtype=np.uint32)
    want = Oracle().durbin_batch(bases, offsets, pairs, at, n_threads=4)
    text = ("# Format = >{RNA sequence id 1},{RNA sequence id 2} {line break} {nucleotide 1}, {nucleotide 2}, "
            "{nucletide matching probability} ...")
    po = want["prob_offsets"] if "prob_offsets" in want else None
    off = 0
    for (a, b) in pairs:
        n, m = len(seqs[a]) + 2, len(seqs[b]) + 2
        mat = want["probs"][off:off + n * m].reshape(n, m)
        off += n * m
        text += f"\n\n>{a},{b}\n"
        for i in range(n):
            for j in range(m):
                if mat[i, j] > 0:
                    text += f"{i - 1},{j - 1},{rust_f32(mat[i, j])} "
    assert out.read_text() == text
