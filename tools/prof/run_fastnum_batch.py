"""FAST_F32 numeric mode on a tRNA-length batch (the FAST build of the batch kernel, csrc/fold_fastnum.cu): 4 096 copies of
the 89-nt bundled tRNA, CONTRAfold, BPP + centroid, three calls.  Profiling driver (see profiles/README.md)."""
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import bench  # noqa: E402
from common import default_tables, load_trnas  # noqa: E402
from rna_algos_b200.api import Handle  # noqa: E402

dev = torch.device("cuda:0"); torch.cuda.set_device(dev)
stream = torch.cuda.Stream()
tt, ct, at = default_tables()
h = Handle(0, tt, ct, at)
seqs = [load_trnas()[5] for _ in range(4096)]
r = bench.FoldRunner(h, seqs, True, [1.0], dev, stream)
h.set_numeric_mode("fast")
with torch.cuda.stream(stream):
    ms = r.timed(2, 1)
print(f"FAST_F32 batch of {r.n} x 89 nt: {ms:.1f} ms -> {r.n / ms * 1e3:.0f} seq/s")
