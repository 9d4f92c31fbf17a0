#!/bin/bash
# ncu evidence of round 2 (ONE gpurun call, one GPU):  bash tools/ncu_r02.sh [tag]
#   launch list of the bench command at n=128 + `--set full` captures of the dominant kernels.
# Graph replay is switched off under ncu (MAMG_GRAPH=0) so that every kernel is a plain launch.
set -u
tag=${1:-r02}
mkdir -p gpurun_out
export MAMG_GRAPH=0
B="python tools/dev_time.py --kind bidomain -n 128 --cycle V --reps 1 --pcg 1"
E="python tools/dev_time.py --kind emi -n 200 --prm default_metric_parameters --cycle V --reps 1 --pcg 1"
$B > gpurun_out/${tag}_plain_b.log 2>&1 || { echo "plain bidomain run failed"; tail -5 gpurun_out/${tag}_plain_b.log; exit 1; }
$E > gpurun_out/${tag}_plain_e.log 2>&1 || { echo "plain emi run failed"; tail -5 gpurun_out/${tag}_plain_e.log; exit 1; }
tail -1 gpurun_out/${tag}_plain_b.log; tail -1 gpurun_out/${tag}_plain_e.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_b.csv $B > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/${tag}_launches_e.csv $E > /dev/null 2>&1
F="ncu --set full --clock-control none --import-source on"
$F -k regex:schwarz_fast -s 100 -c 2 -f -o gpurun_out/${tag}_schwarz_fast $B > /dev/null 2>&1
$F -k regex:sell_gs -s 0 -c 3 -f -o gpurun_out/${tag}_sell_gs $B > /dev/null 2>&1
$F -k regex:"sell_spmv_kernel|agg_sum|sell_scale" -s 0 -c 4 -f -o gpurun_out/${tag}_sell_resid $B > /dev/null 2>&1
$F -k regex:sell_spmv_dot -s 1 -c 2 -f -o gpurun_out/${tag}_sell_spmv_dot $B > /dev/null 2>&1
$F -k regex:schwarz_apply -s 50 -c 2 -f -o gpurun_out/${tag}_schwarz_general $E > /dev/null 2>&1
$F -k regex:"sell_gs|sell_spmv_dot" -s 0 -c 3 -f -o gpurun_out/${tag}_sell_gs_emi $E > /dev/null 2>&1
ls -la gpurun_out/${tag}_*
