"""Parity of the CUDA path at BASELINE.json's own configurations (configs[1..4]) — the shapes, batch, weights and
execution mode (CUDA-graph replay) that bench.py times.  Needs a GPU; the CPU oracle runs on 1-2 clips per case
(tens of seconds), the rest of a batch is covered by batch-permutation invariance."""
import pytest
import torch

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)

FLOW_TOL_PX = 1e-3       # north star: flows within 1e-3 px max-abs (fp32-class modes)
EPE_TOL_PX = 1e-4        # per-clip EPE within 1e-4 px


def build(kind, seed=2, precision=None):
    from accflow_b200.networks import build_flow_estimator
    from accflow_b200.networks.AccFlow_ import AccFlow
    from accflow_b200.weights import make_state_dict
    m = build_flow_estimator(kind)
    if kind.startswith("acc"):
        m = AccFlow(m)
    sd = make_state_dict(kind, seed=seed)
    m.load_state_dict(sd)
    m = m.cuda().eval()
    if precision is not None:
        (m.ofe if kind.startswith("acc") else m).precision = precision
    return m, sd


def maxdiff(a, b):
    return float((a.detach().float().cpu() - b.float()).abs().max())


def epe_triplet(pred, batch, idx):
    from oracle import ops
    bflow, fflow = batch["bflows"][-1][idx:idx + 1], batch["fflows"][-1][idx:idx + 1]
    occ, _ = ops.calc_occ_mask(bflow, fflow)
    return torch.stack(ops.cal_epe(pred, bflow, occ)).reshape(-1)


def test_config1_accflow_raft_9clips_graph_replay():
    """BASELINE configs[1] exactly as bench.py runs it: AccFlow+RAFT, 7 x 512x512, 12 iters/pair, 9 clips per step,
    seed-2 weights, default fp16x2 arithmetic, third call = CUDA-graph replay.  Clips 0 and 8 vs the CPU oracle
    (flows 1e-3 px, per-clip EPE 1e-4 px); the other clips by batch-permutation invariance; replay == eager."""
    from accflow_b200.data import make_batch
    from accflow_b200 import metrics
    from oracle import flow_oracle as fo
    m, sd = build("acc+raft")
    batch = make_batch(list(range(9)), size=512)
    imgs = [t.cuda() for t in batch["imgs"]]
    m.ofe.use_cuda_graph = False
    eager = m(images=imgs, test_mode=False)
    m.ofe.use_cuda_graph = True
    for _ in range(3):                                    # 2 eager warm-ups, then capture + replay
        out = m(images=imgs, test_mode=False)
    out = m(images=imgs, test_mode=False)                 # pure replay
    assert len(out) == 5 and out[0].shape == (9, 2, 512, 512)
    for a, b in zip(out, eager):
        assert maxdiff(a, b.cpu()) == 0.0
    epe_dev = metrics.clip_epe(out[-1], batch["bflows"][-1].cuda(), batch["fflows"][-1].cuda()).cpu()
    for idx in (0, 8):
        ref = fo.accflow_forward(sd, [t[idx:idx + 1] for t in batch["imgs"]])
        for a, b in zip(out, ref):
            assert maxdiff(a[idx:idx + 1], b) < FLOW_TOL_PX, idx
        e_ref = epe_triplet(ref[-1], batch, idx)
        assert float((epe_triplet(out[-1][idx:idx + 1].cpu(), batch, idx) - e_ref).abs().max()) < EPE_TOL_PX
        assert float((epe_dev[idx] - e_ref).abs().max()) < EPE_TOL_PX          # fused metric kernel
    perm = [4, 0, 7, 2, 8, 1, 3, 6, 5]
    out_p = m(images=[t[perm].contiguous() for t in imgs], test_mode=False)
    for a, b in zip(out_p, out):
        assert maxdiff(a, b[perm].cpu()) < 1e-4           # clips are independent (tile order may differ: not bit-exact)


def test_config2_accflow_gma_512():
    """BASELINE configs[2]: AccFlow+GMA over one 7-frame 512x512 clip (4096x4096 attention per pair) vs the oracle."""
    from accflow_b200.data import make_batch
    from oracle import flow_oracle as fo
    m, sd = build("acc+gma")
    batch = make_batch([3], size=512)
    out = m(images=[t.cuda() for t in batch["imgs"]], test_mode=False)
    ref = fo.accflow_forward(sd, batch["imgs"])
    for a, b in zip(out, ref):
        assert maxdiff(a, b) < FLOW_TOL_PX
    assert float((epe_triplet(out[-1].cpu(), batch, 0) - epe_triplet(ref[-1], batch, 0)).abs().max()) < EPE_TOL_PX


def test_config4_raft_1024_32iters_fp16x2_and_bf16():
    """BASELINE configs[4]: 1024x1024, 32 iterations (16384^2 correlation volume).  fp16x2 (default, fp32-class) at
    the 1e-3 px bar; bf16 (the reference's mixed-precision class) with its own STATED tolerance for this config:
    4.0 px max-abs / 0.25 px mean end-point difference against the fp32 oracle (32 recurrent iterations amplify
    the 8-bit-mantissa operand rounding; flows reach tens of px)."""
    from accflow_b200.data import make_clip
    from oracle import flow_oracle as fo
    clip = make_clip(9, size=1024)
    i1, i2 = clip["imgs"][3], clip["imgs"][0]
    m, sd = build("raft")
    ref = fo.flow_estimator(sd, i1, i2, 32)
    out = m(i1.cuda(), i2.cuda(), iters=32)
    assert out.shape == (1, 2, 1024, 1024)
    err = maxdiff(out, ref)
    m.precision = "bf16"
    out16 = m(i1.cuda(), i2.cuda(), iters=32)
    err16 = maxdiff(out16, ref)
    mean16 = float((out16.cpu() - ref).norm(dim=1).mean())
    print(f"1024x1024x32it: fp16x2 max {err:.2e} px; bf16 max {err16:.3f} px mean {mean16:.4f} px; |flow| max {float(ref.abs().max()):.1f}")
    assert err < FLOW_TOL_PX
    assert err16 < 4.0 and mean16 < 0.25


def test_accflow_1024_bf16_two_frames_steps():
    """AccFlow at 1024x1024 (configs[4] names RAFT/AccFlow): one 4-frame clip, 8 iterations, fp16x2 vs oracle."""
    from accflow_b200.data import make_batch
    from oracle import flow_oracle as fo
    m, sd = build("acc+raft")
    m.iters = 8
    batch = make_batch([5], size=1024, frames=4)
    out = m(images=[t.cuda() for t in batch["imgs"]], test_mode=False)
    ref = fo.accflow_forward(sd, batch["imgs"], 8)
    assert len(out) == 2 and out[0].shape == (1, 2, 1024, 1024)
    for a, b in zip(out, ref):
        assert maxdiff(a, b) < FLOW_TOL_PX


@pytest.mark.parametrize("kind", ["acc+raft", "acc+gma"])
def test_warm_start_matches_oracle(kind):
    """Warm-start mode (README TODO; flow_init chaining through raft.py:123-124): same chaining in the oracle."""
    from accflow_b200.data import make_batch
    from oracle import flow_oracle as fo
    m, sd = build(kind)
    m.warm_start, m.warm_iters, m.iters = True, 6, 12
    batch = make_batch([2, 6], size=256, frames=5)
    ref = fo.accflow_forward(sd, batch["imgs"], 12, warm_start=True, warm_iters=6)
    for _ in range(4):                                    # eager, eager, capture, replay
        out = m(images=[t.cuda() for t in batch["imgs"]], test_mode=False)
    assert len(out) == 3
    for a, b in zip(out, ref):
        assert maxdiff(a, b) < FLOW_TOL_PX
    cold = fo.accflow_forward(sd, batch["imgs"], 12)
    assert max(maxdiff(a, b) for a, b in zip(out, cold)) > 1e-3      # the mode really changes the computation
