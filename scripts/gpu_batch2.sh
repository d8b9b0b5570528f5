#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2b}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
timeout 900 python scripts/gru_probe.py > $O/${TAG}_gru_probe.jsonl 2> $O/${TAG}_gru_probe.err; echo "probe rc=$?"
TRACE_PREC=fp16x2 timeout 300 python scripts/mma_trace.py > $O/${TAG}_mma_trace.jsonl 2>&1
timeout 900 python bench.py --ofe gma --clips 4 --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_gma.json 2> $O/${TAG}_bench_gma.err; echo "gma rc=$?"
