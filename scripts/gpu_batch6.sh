#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2f}
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "pair or conv2d or gru or gma" > $O/${TAG}_pytest_pair.log 2>&1; echo "pytest pair rc=$?" >> $O/${TAG}_pytest_pair.log
tail -4 $O/${TAG}_pytest_pair.log
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -4 $O/${TAG}_pytest.log
for PM in 1 0 2; do
  ACCFLOW_TC_PAIR=$PM timeout 900 python bench.py --steps 8 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_pair$PM.json 2> $O/${TAG}_bench_pair$PM.err; echo "bench pair=$PM rc=$?"
  ACCFLOW_TC_PAIR=$PM CLIPS=9 timeout 600 python scripts/conv_breakdown.py > $O/${TAG}_conv_breakdown_pair$PM.txt 2>&1
done
ACCFLOW_TC_PAIR=1 timeout 900 python bench.py --precision fp16 --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_fp16_pair1.json 2> $O/${TAG}_bench_fp16_pair1.err
ACCFLOW_TC_PAIR=2 timeout 900 python bench.py --precision fp16 --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_fp16_pair2.json 2> $O/${TAG}_bench_fp16_pair2.err
