#!/bin/bash
# One 8-GPU box: BASELINE configs[3] (fixed 64-clip sweep on 1/2/4/8 GPUs), configs[4] (1024x1024, 32 iters, bf16) on 8 GPUs,
# and the 2-GPU nn.DataParallel engine-reuse test.
set +e
O=gpurun_out
TAG=${TAG:-r2multi}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $O/${TAG}_gpus.txt
timeout 300 python -m pytest tests/test_gpu_end_to_end.py -m gpu -q -k "dataparallel" > $O/${TAG}_pytest_dp.log 2>&1; echo "dp rc=$?"
for N in 8 4 2; do
  timeout 600 $TR --nproc-per-node $N --master-port $((29500+N)) bench.py --gpus $N --total-clips 64 --steps 4 --warmup 2 > $O/${TAG}_sweep64_${N}gpu.json 2> $O/${TAG}_sweep64_${N}gpu.err; echo "sweep N=$N rc=$?"
done
timeout 600 python bench.py --gpus 1 --total-clips 64 --steps 3 --warmup 2 --no-ref-cuda --no-cpu-baseline > $O/${TAG}_sweep64_1gpu.json 2> $O/${TAG}_sweep64_1gpu.err; echo "sweep N=1 rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29611 bench.py --gpus 8 --size 1024 --iters 32 --precision bf16 --clips 2 --steps 3 --warmup 2 > $O/${TAG}_1024_bf16_8gpu.json 2> $O/${TAG}_1024_bf16_8gpu.err; echo "1024 bf16 rc=$?"
timeout 900 $TR --nproc-per-node 8 --master-port 29612 bench.py --gpus 8 --size 1024 --iters 32 --precision fp16x2 --clips 2 --steps 3 --warmup 2 > $O/${TAG}_1024_fp16x2_8gpu.json 2> $O/${TAG}_1024_fp16x2_8gpu.err; echo "1024 fp16x2 rc=$?"
