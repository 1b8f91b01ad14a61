#!/bin/bash
# A/B runs of bench.py on one GPU: every argument is "label:ENV1=a,ENV2=b" (label alone = defaults).
# Prints label, ms/step, ms per message-phase launch, events/s. Results: gpurun_out/ab_<label>.json
mkdir -p gpurun_out
for spec in "$@"; do
  label="${spec%%:*}"; envs=""
  if [[ "$spec" == *:* ]]; then envs="${spec#*:}"; fi
  envs="${envs//,/ }"
  env $envs python bench.py --steps ${AB_STEPS:-200} --warmup 20 --no-cpu-baseline --no-dse --no-extras ${AB_ARGS} > gpurun_out/ab_${label}.json 2> gpurun_out/ab_${label}.err
  python - "$label" <<'PY'
import json, sys
label = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/ab_{label}.json").read().strip().splitlines()[-1])
    r = d.get("roofline") or {}
    print(f"{label:>14}: {d['ms_per_step']*1e3:8.1f} us/step  fanout {1e3*(r.get('kernel_ms_per_launch') or 0):7.1f} us  {d['value']:.3e} ev/s  e2e {d['e2e']['value']:.3e}")
except Exception as e:
    print(f"{label:>14}: FAILED {e!r}")
    print(open(f"gpurun_out/ab_{label}.err").read()[-600:])
PY
done
