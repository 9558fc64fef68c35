import sys, json
for l in sys.stdin:
    if l.startswith("{"):
        d = json.loads(l)
        r = d.get("roofline") or {}
        print("value %.3e items/s  run %.3f ms | p1 %.3f ms (%.0f GB/s, %.1f%% HBM) p2 %.3f ms | e2e %.3f ms | launches/run %s | clocks %s" % (
            d["value"], d["ms_per_step"], d["phases"]["inner_product_ms"], r.get("achieved", 0), 100 * r.get("frac", 0),
            d["phases"]["multiply_relin_mask_ms"], d["e2e"]["ms_per_step"], d["phases"]["launches_per_run"], d["clocks"]))
        if "serial_ms_per_step" in d["e2e"]: print("e2e serial %.3f ms, pipelined %.3f ms" % (d["e2e"]["serial_ms_per_step"], d["e2e"]["ms_per_step"]))
        ri = d.get("roofline_int")
        if ri: print("phase2: %.3e butterflies/s = %.1f%% of measured butterfly peak %.3e" % (ri["achieved"], 100 * ri["frac"], ri["peak"]))
        if d.get("cpu_baseline"): print("cpu", d["cpu_baseline"])
    else:
        print(l, end="")
