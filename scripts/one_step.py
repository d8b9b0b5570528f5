"""One eager AccFlow+RAFT step bracketed by cudaProfilerStart/Stop, for `ncu --profile-from-start off`.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \\
      --log-file gpurun_out/launches.csv python scripts/one_step.py
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200.data import make_batch
from accflow_b200.networks import build_flow_estimator
from accflow_b200.networks.AccFlow_ import AccFlow
from accflow_b200.weights import make_state_dict

torch.set_grad_enabled(False)
clips = int(os.environ.get("CLIPS", "9"))
kind = os.environ.get("OFE", "raft")
m = AccFlow(build_flow_estimator("acc|" + kind)); m.load_state_dict(make_state_dict("acc+" + kind, seed=2)); m = m.cuda().eval()
m.ofe.precision = os.environ.get("ACCFLOW_PRECISION", "fp16x2"); m.ofe.use_cuda_graph = False
imgs = [t.cuda() for t in make_batch(list(range(clips)), size=int(os.environ.get("SIZE", "512")))["imgs"]]
for _ in range(2):
    m(images=imgs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
m(images=imgs)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
