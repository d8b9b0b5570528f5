#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2d}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
CLIPS=9 timeout 600 python scripts/conv_breakdown.py > $O/${TAG}_conv_breakdown_clips9.txt 2>&1
ACCFLOW_LOOKUP_VEC=0 timeout 900 python bench.py --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_novec.json 2> $O/${TAG}_bench_novec.err; echo "novec rc=$?"
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches_fp16x2.csv python scripts/one_step.py > $O/${TAG}_ncu_launches.log 2>&1
timeout 900 python bench.py --ofe gma --clips 4 --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_gma.json 2> $O/${TAG}_bench_gma.err; echo "gma rc=$?"
