#!/bin/bash
# One gpurun call: GPU tests, default bench, extra bench modes, per-layer breakdown.  Logs under gpurun_out/.
set +e
O=gpurun_out
TAG=${TAG:-r2a}
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/${TAG}_gpu.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
tail -c 600 $O/${TAG}_bench_default.err
CLIPS=9 timeout 600 python scripts/conv_breakdown.py > $O/${TAG}_conv_breakdown_clips9.txt 2>&1
timeout 900 python bench.py --ofe gma --clips 4 --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_gma.json 2> $O/${TAG}_bench_gma.err; echo "gma rc=$?"
