#!/usr/bin/env bash
# A/B of packed-K2 variants selected by environment variables, through bench.py's per-kernel CUDA-event timings.
# Usage (on the GPU box): profiles/tools/k2_ab.sh "MPB_K2_LOCAL=0" "MPB_K2_LOCAL=1" ...
for envs in "$@"; do
  env $envs python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$envs', round(d['ms_per_step'],4), [round(k['ms'],4) for k in d['roofline']['kernels']], d.get('parity_check',{}).get('ok'), d.get('parity_check',{}).get('costs_max_rel_err'))"
done
