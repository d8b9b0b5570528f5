#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2o}
for CFG in "128 592" "64 256" "64 592" "96 256"; do
  set -- $CFG
  export ACCFLOW_TC_BN_CAP=$1 ACCFLOW_TC_MSUB_MIN=$2
  PROBE_PAIRS=18 PROBE_ONLY="256" timeout 300 python scripts/gru_probe.py 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print('bn_cap $1 msub_min $2', d['conv'], d['us_dbg0'], d['us_dbg2'], d['us_dbg4'])
" > $O/${TAG}_probe_bn$1_ms$2.txt
  cat $O/${TAG}_probe_bn$1_ms$2.txt
  timeout 900 python bench.py --steps 6 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_bn$1_ms$2.json 2> $O/${TAG}_bench_bn$1_ms$2.err; echo "bench $CFG rc=$?"
done
export ACCFLOW_TC_BN_CAP=64 ACCFLOW_TC_MSUB_MIN=256
CLIPS=9 timeout 600 python scripts/conv_breakdown.py > $O/${TAG}_conv_breakdown_bn64.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_end_to_end.py -m gpu -x -q > $O/${TAG}_pytest_bn64.log 2>&1; tail -2 $O/${TAG}_pytest_bn64.log
