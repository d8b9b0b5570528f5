#!/bin/bash
# perf experiments: bench.py under the tile-shape knobs (ACCFLOW_TC_BN_CAP / ACCFLOW_TC_MSUB_MIN)
run() { # name, precision, env...
  local name=$1 prec=$2; shift 2
  env "$@" timeout 600 python bench.py --precision $prec --steps 4 --no-cpu-baseline --no-ref-cuda > gpurun_out/${TAG}_$name.json 2> gpurun_out/${TAG}_$name.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value'],1), 'ms', round(d['ms_per_step'],1), d['clocks']['sm_mhz'])"
}
run fp16_default fp16 A=1
run fp16_bn128_ms256 fp16 ACCFLOW_TC_BN_CAP=128 ACCFLOW_TC_MSUB_MIN=256
run fp16_ms256 fp16 ACCFLOW_TC_MSUB_MIN=256
run fp16x2_ms256 fp16x2 ACCFLOW_TC_MSUB_MIN=256
run fp16x2_default fp16x2 A=1
