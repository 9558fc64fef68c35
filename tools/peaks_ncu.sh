#!/bin/bash
# ncu reading of the register-resident butterfly micro-benchmark: what the pipe counters show at the measured ceiling
M=gpu__time_duration.sum,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_alu.sum,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:'peak' --csv --log-file $1 python tools/peaks.py > /dev/null 2>&1
python - $1 <<'PY'
import csv, sys, json, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
agg = collections.OrderedDict()
for r in rows[1:]:
    d = dict(zip(hdr, r))
    agg.setdefault((d["ID"], d["Kernel Name"].split("(")[0]), {})[d["Metric Name"].replace("smsp__average_warps_issue_stalled_","stall_").replace("_per_issue_active.ratio","").split(".")[0]] = float(d["Metric Value"].replace(",", ""))
for k, a in agg.items():
    print(json.dumps({"kernel": k[1], **{m: round(v, 2) for m, v in a.items()}}))
PY
