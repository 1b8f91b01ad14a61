AB_TIMEOUT=80 bash tools/ab_scale.sh 8 ll
grep "bench.py: rank" gpurun_out/sc_ll_n8.err | head -3
