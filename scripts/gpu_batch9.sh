#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2j}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "gru or conv2d" > $O/${TAG}_pytest_k.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest_k.log; tail -3 $O/${TAG}_pytest_k.log
for T in 2 1; do
  ACCFLOW_TC_GRU_TEAMS=$T PROBE_PAIRS=18 timeout 600 python scripts/gru_probe.py > $O/${TAG}_gru_probe_teams$T.jsonl 2> $O/${TAG}_gru_probe_teams$T.err
  ACCFLOW_TC_GRU_TEAMS=$T timeout 900 python bench.py --steps 8 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_teams$T.json 2> $O/${TAG}_bench_teams$T.err; echo "bench teams=$T rc=$?"
  ACCFLOW_TC_GRU_TEAMS=$T timeout 900 python bench.py --precision fp16 --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_fp16_teams$T.json 2> $O/${TAG}_bench_fp16_teams$T.err
done
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
