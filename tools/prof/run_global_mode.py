"""Profiling driver: one launch of the HBM-resident one-CTA mode (mid-length sequences, 300-500 nt, reference-exact)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import default_tables, pack
from rna_algos_b200.api import Handle
tt, ct, at = default_tables()
h = Handle(0, tt, ct, at)
rng = np.random.default_rng(7)
seqs = [rng.integers(0, 4, size=int(L)).astype(np.uint8) for L in rng.integers(420, 500, size=296)]
b, o = pack(seqs)
for _ in range(2):
    r = h.fold_batch(b, o, True, False, [1.0])
print("global-mode batch ok", float(r["logz"][0]))
