"""Regenerate profiles/rNN_* from gpurun_out/ (ncu launch list CSV, ncu full report, bench JSON).
usage: python tools/make_round_profiles.py r01"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def short(name):
    return name.split("(")[0].replace("void ", "").replace("psi::", "")


# --- full-capture summary -------------------------------------------------------------------------------
rep = os.path.join(G, tag + "_run.ncu-rep")
body = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep, "6"], capture_output=True, text=True).stdout
head = ("# Round %s — `ncu --set full --clock-control none --import-source on -k regex:\"k_mac|k_rows|k_cols\" -s 22 -c 11` "
        "on `python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-limb-leg` (`tools/capture_round.sh`)\n\n"
        "The kernels of one `psi_run()` at config B (2^24 vs 2^10, N=16384, L=Lp=4, b=E=47): the inner product over all 47 bins and "
        "the five fused kernels of the FIRST of the two bin groups (23 bins; the second group, 24 bins, is in the report as "
        "launches 7-11 and reads the same within 1-2 points). Report read on the CPU box with `tools/ncu_summary.py`.\n"
        "No tensor-pipe activity anywhere (64-bit residue arithmetic on the integer pipes): `IMAD.WIDE`/`IMAD` run on the "
        "fmaheavy pipe; for what the fmaheavy counter reads at the measured ceiling (92 %%) see `r02_phase2_experiments.md`.\n\n" % tag[1:])
open(os.path.join(P, tag + "_kernels_ncu_full.md"), "w").write(head + body)
ntt_rep = os.path.join(G, tag + "_ntt.ncu-rep")
if os.path.exists(ntt_rep):
    body = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), ntt_rep], capture_output=True, text=True).stdout
    open(os.path.join(P, tag + "_k_ntt_ncu_full.md"), "w").write(
        "# Round %s — `k_ntt` (whole polynomial in shared memory; the offline encode path's only transform, `csrc/ntt.cu`)\n\n"
        "`ncu --set full --clock-control none -k regex:k_ntt -s 4 -c 2` on the bench command: two launches of the device "
        "database build (256 plaintexts per chunk: INTT mod t of the packed slots, then the NTT of every limb).\n\n" % tag[1:] + body)

# --- DRAM traffic per launch ----------------------------------------------------------------------------
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
for r in data:   # one run(): the fused kernels appear once per bin group
    key = short(r[ki]).split("<")[0]
    traffic[key] = traffic.get(key, 0.0) + float(r[ri].replace(",", "")) * scale[units[ri]] + float(r[wi].replace(",", "")) * scale[units[wi]]
json.dump({"2^24_vs_2^10": {"n_gpus": 1, "source": "profiles/%s_kernels_ncu_full.md (ncu --set full, dram__bytes_read.sum + "
                            "dram__bytes_write.sum, one launch)" % tag, "dram_bytes_per_launch": traffic}},
          open(os.path.join(P, tag + "_dram_traffic.json"), "w"), indent=1)

# --- launch list ----------------------------------------------------------------------------------------
src = os.path.join(G, tag + "_launches.csv")
open(os.path.join(P, tag + "_launches_ncu.csv"), "w").write(open(src).read())
rows = [r for r in csv.reader(open(src)) if len(r) > 5]
start = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
hdr = rows[start]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = rows[start + 1:]
seq = []
for i, r in enumerate(data):   # one run() = k_mac_tma followed by the five fused kernels of every bin group
    names = [short(x[ki]) for x in data[i:i + 11]]
    if len(names) == 11 and names[0].startswith("k_mac") and not any(n.startswith("k_mac") for n in names[1:]) \
            and names[10].startswith("k_rows_relin") and (i + 11 == len(data) or short(data[i + 11][ki]).startswith("k_mac")):
        agg = {}
        for n, x in zip(names, data[i:i + 11]):
            agg[n] = agg.get(n, 0.0) + float(x[vi].replace(",", "")) / 1e3
        seq = list(agg.items())
        break
tot = sum(v for _, v in seq)
d = json.load(open(os.path.join(G, "bench_n1.json")))
json.dump(d, open(os.path.join(P, tag + "_bench_n1.json"), "w"))
L = ["# Round %s — launch list of one `psi_run()` (config B: 2^24 server items vs 2^10 client items, 1 B200)\n" % tag[1:],
     "Command (under gpurun, after the same command exited 0 without ncu):",
     "`ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/%s_launches.csv "
     "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-limb-leg`\n" % tag,
     "run() is replayed as a CUDA graph; ncu profiles its kernel nodes one by one: k_mac_tma, then the five fused kernels of "
     "each of the two bin groups (summed per kernel below).\n",
     "Raw CSV: `profiles/%s_launches_ncu.csv`. Per-launch times under ncu are cold-cache and serialised; compare SHARES.\n" % tag,
     "| kernel | ncu time (us) | share of run() |", "|---|---|---|"]
L += ["| %s | %.1f | %.1f %% |" % (n, v, 100 * v / tot) for n, v in seq]
L.append("| total | %.1f | |\n" % tot)
L.append("CUDA-event timing of the same workload in `bench.py` (no profiler, `profiles/%s_bench_n1.json`): run %.3f ms = inner "
         "product %.3f ms (%.1f %% share) + multiply/relinearise/mask %.3f ms; the inner-product share under ncu is %.1f %%.\n"
         % (tag, d["ms_per_step"], d["phases"]["inner_product_ms"], 100 * d["phases"]["inner_product_ms"] / d["ms_per_step"],
            d["phases"]["multiply_relin_mask_ms"], 100 * seq[0][1] / tot))
L.append("Roofline (bench.py, live): k_mac_tma %.0f GB/s of algorithmic traffic = %.1f %% of the measured %.0f GB/s HBM copy "
         "peak; ncu DRAM traffic of that launch %.3f GB vs %.3f GB algorithmic (no re-reads). Phase 2: %.3e butterflies/s = "
         "%.1f %% of the measured register-resident butterfly rate (%.3e/s)."
         % (d["roofline"]["achieved"], 100 * d["roofline"]["frac"], d["roofline"]["peak"], traffic["k_mac_tma"] / 1e9,
            d["roofline"]["algorithmic_bytes_per_launch"] / 1e9, d["roofline_int"]["achieved"], 100 * d["roofline_int"]["frac"],
            d["roofline_int"]["peak"]))
open(os.path.join(P, tag + "_launch_list.md"), "w").write("\n".join(L) + "\n")
# other bench lines of the round (other workloads, N > 1) are copied into profiles/ by hand from the gpurun call that made them
print("\n".join(L))
