#!/bin/bash
# Round evidence on one B200 (run under gpurun): the bench command without a profiler, the ncu launch list of the SAME
# command, one full-set capture of the kernels of one run(), one of the encode path's k_ntt.  usage: tools/capture_round.sh r02
tag=${1:-r02}
G=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-limb-leg --no-nb-leg"
$CMD > $G/${tag}_bench_short.json 2> $G/${tag}_bench_short.err || { echo "bench failed"; tail -5 $G/${tag}_bench_short.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $G/${tag}_launches.csv $CMD > $G/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_mac|k_rows|k_cols' -s 22 -c 11 -f -o $G/${tag}_run $CMD > $G/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_ntt' -s 4 -c 2 -f -o $G/${tag}_ntt $CMD > $G/${tag}_ncu_ntt.log 2>&1
python bench.py > $G/bench_n1.json 2> $G/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $G/bench_ref.json 2> $G/bench_ref.err
ls -la $G/${tag}_* | tail -12
tail -c 600 $G/bench_ref.json
