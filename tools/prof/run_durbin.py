"""Profiling driver: the Durbin kernel on the 15 tRNA pairs tiled to 30720 pairs."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import default_tables, load_trnas, pack
from rna_algos_b200.api import Handle
tt, ct, at = default_tables()
h = Handle(0, None, None, at)
b, o = pack(load_trnas())
base = [(a, c) for a in range(6) for c in range(a + 1, 6)]
pairs = np.array([base[i % 15] for i in range(30720)], dtype=np.uint32)
for _ in range(2):
    r = h.durbin_batch(b, o, pairs)
print("durbin ok", float(r["probs"].sum()))
