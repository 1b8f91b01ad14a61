python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash tools/ab_bench.sh s6def soma4:SFE_LIB_PATH=sana-fe_b200/variants/soma4/libsanafe_b200.so
bash tools/ncu_capture.sh r2b > gpurun_out/ncu_capture.log 2>&1; tail -3 gpurun_out/ncu_capture.log
