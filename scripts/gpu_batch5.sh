#!/bin/bash
set +e
O=gpurun_out
TAG=${TAG:-r2e}
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off"
timeout 900 $NCU -o $O/${TAG}_conv_targets python scripts/ncu_targets.py gru_zr gru_q convc2 enc1 convc1 > $O/${TAG}_ncu_conv.log 2>&1; echo "ncu conv rc=$?"
timeout 900 $NCU -o $O/${TAG}_aux_targets python scripts/ncu_targets.py lookup_vec lookup_fast corr_gemm stem attn agg > $O/${TAG}_ncu_aux.log 2>&1; echo "ncu aux rc=$?"
ls -la $O/*.ncu-rep
timeout 900 python bench.py --precision fp16 --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_fp16.json 2> $O/${TAG}_bench_fp16.err; echo "fp16 rc=$?"
timeout 900 python bench.py --precision bf16 --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_bf16.json 2> $O/${TAG}_bench_bf16.err; echo "bf16 rc=$?"
timeout 900 python bench.py --total-clips 64 --steps 2 --warmup 1 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_bench_sweep64_1gpu.json 2> $O/${TAG}_bench_sweep64_1gpu.err; echo "sweep rc=$?"
timeout 900 python bench.py --warm-start --steps 5 --warmup 3 --no-ref-cuda > $O/${TAG}_bench_warm.json 2> $O/${TAG}_bench_warm.err; echo "warm rc=$?"
