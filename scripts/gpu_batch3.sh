#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2c}
timeout 1500 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${TAG}_pytest.log
tail -5 $O/${TAG}_pytest.log
PROBE_PAIRS=27 timeout 900 python scripts/gru_probe.py > $O/${TAG}_gru_probe.jsonl 2> $O/${TAG}_gru_probe.err; echo "probe rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_default.json 2> $O/${TAG}_bench_default.err; echo "bench rc=$?"
CLIPS=9 timeout 600 python scripts/conv_breakdown.py > $O/${TAG}_conv_breakdown_clips9.txt 2>&1
timeout 900 python bench.py --ofe gma --clips 4 --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_gma.json 2> $O/${TAG}_bench_gma.err; echo "gma rc=$?"
timeout 900 python bench.py --precision fp16 --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_fp16.json 2> $O/${TAG}_bench_fp16.err; echo "fp16 rc=$?"
