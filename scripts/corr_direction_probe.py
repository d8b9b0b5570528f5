import os, sys, json, math
sys.path.insert(0, os.getcwd())
import torch
from accflow_b200 import _lib as L
from accflow_b200.engine import FlowEstimatorEngine, Kernels, View
torch.set_grad_enabled(False)
K = Kernels(torch.device("cuda:0"), "fp16x2")
B, h, w, D = 18, 64, 64, 256
f1 = View(torch.randn(B, h, w, D, device="cuda")); f2 = View(torch.randn(B, h, w, D, device="cuda"))
eng = FlowEstimatorEngine.__new__(FlowEstimatorEngine); eng.k = K
for _ in range(3): eng.corr_pyramid(f1, f2, "t")
torch.cuda.synchronize()
ts = []
for i in range(8):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); eng.corr_pyramid(f1, f2, "t"); b.record(); torch.cuda.synchronize()
    ts.append(round(a.elapsed_time(b) * 1e3, 1))
print(json.dumps({"serpentine": os.environ.get("ACCFLOW_TC_SERPENTINE", "1"), "us_per_call(gemm+split+pool)": ts}))
