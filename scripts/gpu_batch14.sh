#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2p}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log; tail -3 $O/${TAG}_pytest.log
TRACE_PREC=fp16x2 timeout 300 python scripts/mma_trace.py > $O/${TAG}_mma_trace.jsonl 2>&1
TRACE_PREC=fp16 timeout 300 python scripts/mma_trace.py > $O/${TAG}_mma_trace_fp16.jsonl 2>&1
timeout 900 python bench.py --steps 8 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
timeout 900 python bench.py --precision fp16 --steps 5 --warmup 3 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_fp16.json 2> $O/${TAG}_bench_fp16.err; echo "fp16 rc=$?"
