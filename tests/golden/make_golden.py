"""Regenerates tests/golden/trna_oracle.npz: oracle outputs for the 6 bundled tRNAs (both models, the reference
binary's threshold sweep 2^-7..2^10 plus gamma = 1).  The reference itself (Rust) cannot run in this environment and its
tests hold no golden values (tests/tests.rs:33,38,74), so these vectors pin the ORACLE (regression) — see DESIGN.md §5.
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from common import default_tables, load_trnas, pack  # noqa: E402
from oracle_lib import Oracle  # noqa: E402

GAMMAS = [1.0] + [float(np.float32(2.0) ** p) for p in range(-7, 11)]


def main():
    tt, ct, at = default_tables()
    orc = Oracle()
    seqs = load_trnas()
    bases, offsets = pack(seqs)
    out = {"gammas": np.array(GAMMAS, dtype=np.float32)}
    for contra in (False, True):
        r = orc.fold_batch(bases, offsets, contra, False, tt, ct, GAMMAS, n_threads=1)
        k = "contra" if contra else "turner"
        out[k + "_logz"] = r["logz"]
        out[k + "_bpp"] = r["bpp"]
        out[k + "_structs"] = r["structs"]
        out[k + "_expect_acc"] = r["expect_acc"]
    pairs = np.array([(a, b) for a in range(6) for b in range(a + 1, 6)], dtype=np.uint32)
    d = orc.durbin_batch(bases, offsets, pairs, at, n_threads=1)
    out["durbin_probs"] = d["probs"]
    np.savez_compressed(os.path.join(HERE, "trna_oracle.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
