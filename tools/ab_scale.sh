#!/bin/bash
# A/B runs of the partitioned bench on N GPUs: tools/ab_scale.sh N label:ENV=a,ENV2=b ...
N=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  label="${spec%%:*}"; envs=""
  if [[ "$spec" == *:* ]]; then envs="${spec#*:}"; fi
  envs="${envs//,/ }"
  env $envs timeout ${AB_TIMEOUT:-100} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps ${AB_STEPS:-200} --warmup 20 > gpurun_out/sc_${label}_n$N.json 2> gpurun_out/sc_${label}_n$N.err
  python - "$label" "$N" <<'PY'
import json, sys
label, n = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(f"gpurun_out/sc_{label}_n{n}.json").read().strip().splitlines()[-1])
    print(f"{label:>14} N={n}: {d['ms_per_step']*1e3:8.1f} us/step  {d['value']:.3e} ev/s  sha {d.get('raster_sha','')[:12]} equal={d.get('raster_sha_equal_on_all_ranks')}")
except Exception as e:
    print(f"{label:>14} N={n}: FAILED {e!r}")
    print(open(f"gpurun_out/sc_{label}_n{n}.err").read()[-800:])
PY
done
