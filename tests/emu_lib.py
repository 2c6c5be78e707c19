"""ctypes binding of the host emulator of the v2 fold kernel (tests/emu/emu_fold.cpp).  Test infrastructure only."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from rna_algos_b200.tables import ContraTables, TurnerTables

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "emu", "emu_fold.cpp")
LIB = os.path.join(HERE, "emu", "_emu.so")
CSRC = os.path.join(os.path.dirname(HERE), "rna_algos_b200", "csrc")


def build():
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", LIB, SRC],
                   check=True, capture_output=True)


class Emu:
    def __init__(self):
        build()
        self.lib = C.CDLL(LIB)
        self.lib.emu_fold.restype = C.c_int
        self.lib.emu_fold.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(TurnerTables),
                                      C.POINTER(ContraTables), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                      C.POINTER(C.c_float)]

    def fold(self, seq, contra, allows_short, tt, ct, nX=32, nY=64, nZ=64, order=0, tcap=-1):
        """tcap: -1 = term streams (unbounded), 0 = scores on the fly, n = streams if they fit n terms."""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        L = seq.shape[0]
        bpp = np.empty(L * (L - 1) // 2, dtype=np.float32)
        logz = C.c_float()
        rc = self.lib.emu_fold(seq.ctypes.data, L, int(contra), int(allows_short), C.byref(tt), C.byref(ct), nX, nY, nZ,
                               order, tcap, bpp.ctypes.data, C.byref(logz))
        assert rc == 0
        return bpp, np.float32(logz.value)
