"""Register-resident butterfly rate against resident warps per scheduler (psi_bench_pipe_peak, kind bits 4..7 = blocks per SM)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import psi_b200 as P
for per_sm in (3, 8):
    row = {"blocks_per_sm": per_sm, "warps_per_scheduler": per_sm * 2}
    for kind, name in ((1, "butterflies_per_s"), (4, "exchanges_smem_per_s"), (5, "exchanges_shfl_per_s")):
        v = ctypes.c_double()
        P.capi.check(P.lib().psi_bench_pipe_peak(0, kind | (per_sm << 4), ctypes.byref(v)))
        row[name] = float("%.4g" % v.value)
    print(json.dumps(row), flush=True)
