"""GPU: per-layer-shape breakdown of the conv/GEMM launches of one AccFlow+RAFT step (eager, CUDA events)."""
import collections, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from accflow_b200.data import make_batch
from accflow_b200.networks import build_flow_estimator
from accflow_b200.networks.AccFlow_ import AccFlow
from accflow_b200.weights import make_state_dict
import accflow_b200.engine as E

torch.set_grad_enabled(False)
prec = os.environ.get("ACCFLOW_PRECISION", "fp16x2")
clips = int(os.environ.get("CLIPS", "4"))
m = AccFlow(build_flow_estimator("acc|raft")); m.load_state_dict(make_state_dict("acc+raft", seed=2)); m = m.cuda().eval()
m.ofe.precision = prec; m.ofe.use_cuda_graph = False
imgs = [t.cuda() for t in make_batch(list(range(clips)), size=512)["imgs"]]
for _ in range(2): m(images=imgs)
eng = m.engine(torch.device("cuda:0"))
# tag each conv launch with its shape
orig = E.Kernels.conv
tags = []
def conv(self, pc, srcs, *a, **kw):
    s0 = srcs[0]
    tags.append((pc.kh, pc.kw, pc.stride, sum(s.c for s in srcs), kw.get("cout") or pc.cout, s0.b * s0.h * s0.w))
    return orig(self, pc, srcs, *a, **kw)
E.Kernels.conv = conv
eng.k.profile = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); m(images=imgs); e1.record(); torch.cuda.synchronize()
tot = e0.elapsed_time(e1)
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for tag, (a, b, fl) in zip(tags, eng.k.profile):
    ms = a.elapsed_time(b)
    agg[tag][0] += 1; agg[tag][1] += ms; agg[tag][2] += fl
cm = sum(v[1] for v in agg.values())
print(f"precision {prec}: step {tot:.1f} ms, conv launches {len(tags)}, conv time {cm:.1f} ms")
for tag, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:int(os.environ.get("TOP", "60"))]:
    kh, kw, st, cin, cout, pix = tag
    print(f"{v[1]:7.2f} ms {100*v[1]/cm:5.1f}%  n={v[0]:4d}  avg {1e3*v[1]/v[0]:7.1f} us  {v[2]/v[1]/1e9:7.1f} TFLOP/s   {kh}x{kw}/s{st} {cin}->{cout}  pix={pix}")
