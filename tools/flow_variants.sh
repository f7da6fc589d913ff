#!/bin/bash
# Sweeps the launch shapes of the fixed-step flow kernel (PB_SRKN_VARIANT,
# PB_FLOW_GROUPS, PB_FLOW_SUB_STEPS) on the headline workload; prints value
# (probe-stages/s), kernel ms per 100 steps of 10^5 probes, e2e value.
for cfg in "PB_SRKN_VARIANT=0" "PB_SRKN_VARIANT=3" "PB_SRKN_VARIANT=5" "PB_SRKN_VARIANT=5 PB_FLOW_GROUPS=8" "PB_SRKN_VARIANT=5 PB_FLOW_SUB_STEPS=5"; do
  echo "== $cfg"
  env $cfg python bench.py --steps 6 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['kernel_ms'], d['e2e']['value'])"
done
