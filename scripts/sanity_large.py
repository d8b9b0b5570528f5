"""GPU sanity at BASELINE configs[2] / configs[4] shapes: AccFlow+GMA 512^2 clip and RAFT 1024^2 / 32 iters."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200.data import make_batch
from accflow_b200.networks import build_flow_estimator
from accflow_b200.networks.AccFlow_ import AccFlow
from accflow_b200.weights import make_state_dict
torch.set_grad_enabled(False)

def timeit(fn, n=3):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t=time.time()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.time()-t)/n

for prec in os.environ.get("SANITY_PREC", "fp16x2,bf16").split(","):
    m = AccFlow(build_flow_estimator("acc|gma")); m.load_state_dict(make_state_dict("acc+gma", seed=2)); m = m.cuda().eval(); m.ofe.precision = prec
    b = make_batch([0, 1], size=512)
    imgs = [t.cuda() for t in b["imgs"]]
    dt = timeit(lambda: m(images=imgs))
    print(json.dumps({"case": "AccFlow+GMA 2 clips x 7 x 512x512", "precision": prec, "ms_per_step": dt*1e3, "flows_per_s": 10/dt,
                      "mem_GB": torch.cuda.max_memory_allocated()/1e9}))
    del m
    torch.cuda.empty_cache()
    r = build_flow_estimator("raft"); r.load_state_dict(make_state_dict("raft", seed=1)); r = r.cuda().eval(); r.precision = prec
    b = make_batch([3], size=1024)
    i1, i2 = b["imgs"][3].cuda(), b["imgs"][0].cuda()
    dt = timeit(lambda: r(i1, i2, iters=32))
    out = r(i1, i2, iters=32)
    print(json.dumps({"case": "RAFT pair 1024x1024, 32 iters", "precision": prec, "ms_per_pair": dt*1e3, "finite": bool(torch.isfinite(out).all()),
                      "mem_GB": torch.cuda.max_memory_allocated()/1e9}))
    del r
    torch.cuda.empty_cache()
