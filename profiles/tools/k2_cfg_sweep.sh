#!/usr/bin/env bash
# A/B of the packed K2 launch shapes (MPB_K2_CFG) through bench.py's per-kernel CUDA-event timings.  Run on the GPU box.
for cfg in 8x2 8x3 6x3 4x5 4x6; do
  MPB_K2_CFG=$cfg python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-other-configs --no-parity-check 2>/dev/null |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$cfg', round(d['ms_per_step'],4), [round(k['ms'],4) for k in d['roofline']['kernels']])"
done
