import ctypes, sys, os
sys.path.insert(0, os.getcwd())
import psi_b200 as P
for kind, name in ((1, "shoup butterflies/s (compiler)"), (2, "hand-scheduled"), (6, "FP64-quotient butterflies/s"), (7, "FP64-quotient range violations"), (1 + (3 << 4), "shoup @3 blocks/SM"), (6 + (3 << 4), "FP64 @3 blocks/SM")):
    v = ctypes.c_double()
    rc = P.lib().psi_bench_pipe_peak(0, kind, ctypes.byref(v))
    print(name, rc, '%.4g' % v.value)
