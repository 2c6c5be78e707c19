"""Runs the host emulator of the kernel schedules under AddressSanitizer (subprocess of tests/test_emu_parity.py).
The emulator executes the product's phase functions on exactly-sized host buffers, so an out-of-bounds read or write
of any phase — including the running-pointer loops of the cooperative kernel — aborts here."""
import sys

sys.path.insert(0, sys.argv[1])
sys.path.insert(0, sys.argv[1] + "/tests")
import emu_lib

emu_lib.LIB = sys.argv[2]
emu_lib.build = lambda: None
from common import default_tables, load_trnas, random_seqs
from emu_lib import Emu

tt, ct, _ = default_tables()
e = Emu()
for contra in (False, True):
    for s in load_trnas()[:2] + random_seqs(9, [1, 2, 3, 4, 5, 6, 7, 17, 33, 40, 64, 65, 97, 131]):
        for kw in (dict(order=3), dict(order=3, nX=5, nY=3, nZ=40), dict(order=3, tcap=0), dict(order=2), dict(order=0),
                   dict(order=1, tcap=1000)):
            e.fold(s, contra, False, tt, ct, **kw)
print("asan-clean")
