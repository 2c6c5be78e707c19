"""Profiling driver: FAST numeric mode (f32), one 2048-nt sequence on the cooperative grid."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import default_tables, pack
from rna_algos_b200.api import Handle
tt, ct, at = default_tables()
h = Handle(0, tt, ct, at)
h.set_numeric_mode("fast")
seq = np.random.default_rng(2).integers(0, 4, size=2048).astype(np.uint8)
b, o = pack([seq])
for _ in range(2):
    r = h.fold_batch(b, o, True, False, [])
print("fast long ok", float(r["logz"][0]))
