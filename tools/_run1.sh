timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/final_n8.json 2> gpurun_out/final_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/final_n8.json").read().strip().splitlines()[-1])
print(d["n_gpus"], d["ms_per_step"], d["value"], d["raster_sha"][:12], d["raster_sha_equal_on_all_ranks"], d["dse"])
PY
grep -v "^W1\|OMP_NUM\|^\*\*\*" gpurun_out/final_n8.err | tail -5
