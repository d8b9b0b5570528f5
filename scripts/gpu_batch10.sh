#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2k}
timeout 300 python scripts/corr_bench.py > $O/${TAG}_corr_bench.jsonl 2> $O/${TAG}_corr_bench.err; echo "corr rc=$?"; cat $O/${TAG}_corr_bench.jsonl
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
timeout 900 python bench.py --steps 8 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
