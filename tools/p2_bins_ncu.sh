M=gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,launch__grid_size
for b in 6 12 47; do
ncu --metrics $M --clock-control none -k regex:'k_rows|k_cols' -s 5 -c 5 --csv --log-file gpurun_out/r2_p2b_$b.csv python tools/p2_probe.py $b 1 --once > /dev/null 2>&1
python - gpurun_out/r2_p2b_$b.csv $b <<'PY'
import csv, sys, json, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
agg = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "").split("<")[0]
    agg.setdefault(name, {})[d["Metric Name"].split(".")[0]] = float(d["Metric Value"].replace(",", ""))
print(json.dumps({"b": int(sys.argv[2]), "kernels": {k: [round(a["gpu__time_duration"]/1000,1), a["launch__grid_size"], round(a["sm__pipe_fmaheavy_cycles_active"],1)] for k, a in agg.items()}}))
PY
done
