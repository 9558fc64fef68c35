#!/bin/bash
# usage: tools/p2_variants.sh out_prefix lib1 lib2 ...   (libs relative to nested-hashing-psi_b200/, "default" = the product build)
# For each library: event-timed phase 2 (1 and 2 bin groups) and an ncu per-launch list of one evaluation with one bin group.
out=$1; shift
M=gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_lsu.sum
for lib in "$@"; do
  if [ "$lib" = default ]; then unset PSI_B200_LIB; else export PSI_B200_LIB=$PWD/nested-hashing-psi_b200/$lib; fi
  python tools/p2_probe.py 47 1 >> $out.jsonl 2>> $out.err
  python tools/p2_probe.py 47 2 >> $out.jsonl 2>> $out.err
  ncu --metrics $M --clock-control none -k regex:'k_rows|k_cols' -s 5 -c 5 --csv --log-file $out.$lib.csv python tools/p2_probe.py 47 1 --once > /dev/null 2>> $out.err
  python - "$out.$lib.csv" "$lib" >> $out.jsonl <<'PY'
import csv, sys, json, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
agg = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    name = d["Kernel Name"].split("(")[0].replace("void ", "")
    agg.setdefault(name, {})[d["Metric Name"].split(".")[0]] = float(d["Metric Value"].replace(",", ""))
print(json.dumps({"lib": sys.argv[2], "kernels": {k: {m: round(v, 1) for m, v in a.items()} for k, a in agg.items()}}))
PY
done
cat $out.jsonl
