"""Indicative timing of the reference ALGORITHM in eager PyTorch on the GPU (not a test; not collected).

The reference sources cannot travel to the GPU box, so the reference's own CUDA path cannot be
timed there.  The oracle port executes the same sequence of ATen ops (conv2d, bmm/einsum,
avg-pool, gather-based sampling, softmax), so running it on CUDA gives the order of magnitude of
"stock eager PyTorch on a B200" for this workload: fp32 (cudnn TF32 allowed, torch default) and
under fp16 autocast (the reference's default, networks/__init__.py:8).

    python tests/perf_eager_port.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200.data import make_batch
from accflow_b200.weights import make_state_dict
from oracle import flow_oracle as fo

torch.set_grad_enabled(False)
torch.backends.cudnn.benchmark = True                      # test_cvo.py:115
dev = torch.device("cuda:0")
sd = {k: v.to(dev) for k, v in make_state_dict("acc+raft", seed=2).items()}
for clips in (1, 4):
    imgs = [t.to(dev) for t in make_batch(list(range(clips)), size=512)["imgs"]]
    for mode in ("fp32", "fp16-autocast"):
        def run():
            if mode == "fp32":
                return fo.accflow_forward(sd, imgs)
            with torch.autocast("cuda", dtype=torch.float16):
                return fo.accflow_forward(sd, imgs)
        for _ in range(2):
            run()
        torch.cuda.synchronize()
        t0 = time.time()
        n = 3
        for _ in range(n):
            run()
        torch.cuda.synchronize()
        dt = (time.time() - t0) / n
        print(json.dumps({"impl": "oracle port, eager PyTorch CUDA", "mode": mode, "clips": clips, "ms_per_step": dt * 1e3,
                          "flows_per_s": 5 * clips / dt}))
