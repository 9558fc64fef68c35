import ctypes, sys
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import psi_b200 as P
for kind, name in ((0, "IMAD.WIDE.U32/s"), (1, "shoup butterflies/s (compiler)"), (2, "shoup butterflies/s (hand-scheduled)"), (3, "shoup butterflies/s (PTX block)")):
    v = ctypes.c_double()
    rc = P.lib().psi_bench_pipe_peak(0, kind, ctypes.byref(v))
    print(name, rc, '%.4g' % v.value)
