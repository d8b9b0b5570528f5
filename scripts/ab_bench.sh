#!/bin/bash
# usage: scripts/ab_bench.sh <out.jsonl> "<env A>" "<env B>" [bench args...]
# Alternates two environments (A B A B) through the fast bench on ONE box, so that box-to-box clock differences
# (sw_power_cap: 1.57-1.79 GHz) cancel.  An environment may carry per-arm bench arguments as BENCH_EXTRA=--clips=6.  Each line: {"env", "value", "ms_per_step", "sm_mhz"}.
OUT=$1; A=$2; B=$3; shift 3
for rep in 1 2; do
  for E in "$A" "$B"; do
    env $E bash -c 'python bench.py --no-cpu-baseline --no-ref-cuda $BENCH_EXTRA "$@"' ab "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print(json.dumps({'env': '''$E''', 'value': round(d['value'],1), 'ms_per_step': round(d['ms_per_step'],2), 'e2e': round(d['e2e']['value'],1), 'sm_mhz': d['clocks']['sm_mhz'], 'conv_ms_eager': round(d['roofline']['ms_in_step'],2)}))" >> $OUT
  done
done
