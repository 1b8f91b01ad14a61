SFE_LIB_PATH=sana-fe_b200/variants/tl/libsanafe_b200.so SFE_TIMELINE=1 bash tools/ab_scale.sh 8 tl
python tools/timeline_partitioned.py "gpurun_out/timeline_n8_r[03].npy"
