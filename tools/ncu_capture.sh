#!/bin/bash
# ncu evidence of the step kernels on the bench workload (one B200): launch list + `--set full` of the message-phase and
# neuron-phase kernels. Usage: tools/ncu_capture.sh <tag>   -> gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_{fanout,soma}.ncu-rep
tag=${1:-r2}
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dse --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv $CMD > gpurun_out/${tag}_ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fanout_kernel -s 4 -c 2 -o gpurun_out/${tag}_fanout $CMD > gpurun_out/${tag}_ncu_fanout.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:soma_kernel -s 4 -c 2 -o gpurun_out/${tag}_soma $CMD > gpurun_out/${tag}_ncu_soma.log 2>&1
ls -la gpurun_out/${tag}_*
