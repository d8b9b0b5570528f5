#!/usr/bin/env python
"""Benchmark of the AccFlow hot path (BASELINE.json: long-range flow pairs/sec, 7-frame
512x512 clips, 12 GRU iterations per pair).

  python bench.py --gpus 1 --steps 10 --warmup 3            # this implementation (BASELINE configs[1])
  python bench.py --ofe gma                                  # configs[2]: AccFlow+GMA
  python bench.py --total-clips 64                           # configs[3]: fixed 64-clip sweep (strong scaling)
  python bench.py --size 1024 --iters 32 --precision bf16 --clips 2      # configs[4]
  python bench.py --impl reference --steps 1 --warmup 0     # reference algorithm on the host CPU cores
  python bench.py --impl reference-cuda                      # the reference's schedule in eager PyTorch on the GPU
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of AccFlow backward accumulation over `--clips` synthetic CVO-shaped clips per GPU
(5 long-range flows per 7-frame clip); with `--total-clips T` a step is the whole sweep of T clips dealt
round-robin to the ranks in micro-batches of `--clips`.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

FRAMES = 7
FLOWS_PER_CLIP = FRAMES - 2
# conv/GEMM FLOPs (2*MAC) necessary for identical output, B=1, 512x512, 12 iters.
NECESSARY_GFLOP_PER_CLIP_512 = 3984.5
FLOW_TOL_PX, EPE_TOL_PX = 1e-3, 1e-4          # north star: fp32-class flows / per-clip EPE


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "reference-cuda"])
    ap.add_argument("--clips", type=int, default=9,
                    help="clips per GPU per (micro-)step (9 clips = 18/27 pairs fill the 148 SMs with 7.8/11.7 tile rounds)")
    ap.add_argument("--total-clips", type=int, default=0,
                    help="BASELINE configs[3]: a step is a sweep over this many clips, sharded over the ranks (strong scaling)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--ofe", default="raft", choices=["raft", "gma"])
    ap.add_argument("--precision", default=os.environ.get("ACCFLOW_PRECISION", "fp16x2"),
                    choices=["fp32", "bf16x3", "fp16x2", "bf16", "fp16"],
                    help="conv/GEMM arithmetic: fp16x2 / bf16x3 = tcgen05 split products (fp32-class, parity-gated at 1e-3 px)")
    ap.add_argument("--warm-start", action="store_true", help="AccFlow.warm_start (README TODO; raft.py:123-124 flow_init chaining)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (parity + cpu_baseline)")
    ap.add_argument("--no-ref-cuda", action="store_true", help="skip the eager-PyTorch-CUDA proxy of the reference")
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    fallback = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            d.setdefault("bf16_tflops", fallback["bf16_tflops"])
            d.setdefault("bf16_tflops_sustained", d["bf16_tflops"])     # timed inside a long step: sustained figure
            d.setdefault("hbm_gbs", fallback["hbm_gbs"])
            return d, "measured"
        except (OSError, ValueError):
            pass
    return fallback, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 8 for n, v in zip(names, r[4:8]) if v.lower().startswith("active")})
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": reasons}


def make_inputs(clip_ids, size):
    from accflow_b200.data import make_batch
    return make_batch(clip_ids, size=size, frames=FRAMES)


def workload_config(args, clips):
    which = "configs[1]" if args.ofe == "raft" else "configs[2]"
    if args.total_clips:
        which = "configs[3]"
    if args.size == 1024:
        which = "configs[4]"
    cfg = {"workload": f"AccFlow+{args.ofe.upper()} backward accumulation, {FRAMES}-frame {args.size}x{args.size} CVO-shaped synthetic clip, "
                       f"{args.iters} iters/pair (BASELINE {which})",
           "clips_per_gpu_per_step": clips, "frames": FRAMES, "size": args.size, "iters": args.iters,
           "weights": "seeded random (accflow_b200.weights, seed 2)",
           "l2_policy": "working set per step (correlation volumes + activations) exceeds the 126 MB L2"}
    if args.total_clips:
        cfg["total_clips_per_step"] = args.total_clips
    if args.warm_start:
        cfg["warm_start"] = True
    return cfg


# ------------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference algorithm (oracle port of the reference's fp32 CPU path) on host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from accflow_b200.weights import make_state_dict
    from oracle import flow_oracle as fo
    torch.set_grad_enabled(False)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(f"acc+{args.ofe}", seed=2)
    batch = make_inputs([0], args.size)
    budget_s = 150.0
    t_begin = time.time()
    for _ in range(min(args.warmup, 1)):
        fo.accflow_forward(sd, batch["imgs"], args.iters)
    times = []
    for _ in range(max(1, args.steps)):
        t0 = time.time()
        fo.accflow_forward(sd, batch["imgs"], args.iters)
        times.append(time.time() - t0)
        if time.time() - t_begin > budget_s:
            break
    t = statistics.mean(times)
    value = FLOWS_PER_CLIP / t
    sample = f"1 clip x {FRAMES} frames {args.size}x{args.size}, {args.iters} iters/pair per step, {len(times)} steps (time-boxed {budget_s:.0f}s)"
    line = {"impl": "reference", "metric": "long-range flow pairs/sec", "value": value, "unit": "flows/s",
            "n_gpus": args.gpus, "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * t,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, clips=1),
            "cpu_baseline": {"value": value, "unit": "flows/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "flows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- reference schedule on the GPU
def ref_cuda_proxy(args, dev, clips_list=(1, 4, 9, 10), ours_clip0=None, budget_s=75.0):
    """North-star denominator: the reference's own PyTorch-CUDA path.  The reference (Python) cannot travel to
    the GPU box, so `oracle/eager_ref.py` restates its execution schedule with the same library calls
    (tests/test_eager_ref_faithful.py: identical outputs and ATen-op histogram vs the unmodified reference).
    Timed like test_cvo.py runs it: cudnn.benchmark=True, fp16 autocast default, batch 10 (test_cvo.py:114-115)."""
    from accflow_b200.weights import make_state_dict
    from oracle import eager_ref as er
    sd = {k: v.to(dev) for k, v in make_state_dict(f"acc+{args.ofe}", seed=2).items()}
    out = {"what": "oracle/eager_ref.py: the reference's schedule (same ATen/cuDNN/torchvision calls, dead work "
                   "included) in eager PyTorch on this GPU; the reference sources cannot travel to the GPU box",
           "runs": []}
    t_begin = time.time()
    fp32_clip0 = None
    for mode in ("fp16_autocast", "fp32_tf32off"):
        er.configure_like_test_cvo(fp32_exact=(mode == "fp32_tf32off"))
        for b in clips_list:
            if time.time() - t_begin > budget_s and out["runs"]:
                break
            imgs = [t.to(dev) for t in make_inputs(list(range(b)), args.size)["imgs"]]
            run = lambda: er.accflow_forward(sd, imgs, args.iters, mixed_precision=(mode == "fp16_autocast"))
            try:
                for _ in range(2):
                    flows = run()                                  # cudnn.benchmark autotune + allocator warm-up
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                n = 2
                e0.record()
                for _ in range(n):
                    flows = run()
                e1.record()
                torch.cuda.synchronize()
            except torch.OutOfMemoryError:
                out["runs"].append({"mode": mode, "clips": b, "error": "out of memory"})
                break
            ms = e0.elapsed_time(e1) / n
            out["runs"].append({"mode": mode, "clips": b, "ms_per_step": ms, "flows_per_s": FLOWS_PER_CLIP * b / (ms / 1e3)})
            if mode == "fp32_tf32off" and fp32_clip0 is None:
                fp32_clip0 = [f[0:1].float().cpu() for f in flows]
            if mode == "fp16_autocast" and b == clips_list[0]:
                out["_fp16_clip0"] = [f[0:1].float().cpu() for f in flows]
            del imgs, flows
            torch.cuda.empty_cache()
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = True
    best = [r for r in out["runs"] if r.get("mode") == "fp16_autocast" and "flows_per_s" in r]
    if best:
        top = max(best, key=lambda r: r["flows_per_s"])
        out["value"], out["unit"], out["at_clips"] = top["flows_per_s"], "flows/s", top["clips"]
    fp16_clip0 = out.pop("_fp16_clip0", None)
    if fp32_clip0 is not None:
        md = lambda a, b: max(float((x - y).abs().max()) for x, y in zip(a, b))
        if fp16_clip0 is not None:
            out["ref_fp16_autocast_vs_ref_fp32_max_px"] = md(fp16_clip0, fp32_clip0)
        if ours_clip0 is not None:
            out["ours_vs_ref_fp32_max_px"] = md(ours_clip0, fp32_clip0)
    return out


def run_reference_cuda(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_grad_enabled(False)
    assert torch.cuda.is_available(), "--impl reference-cuda needs a CUDA device"
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    r = ref_cuda_proxy(args, dev, budget_s=240.0)
    line = {"impl": "reference-cuda", "metric": "long-range flow pairs/sec", "value": r.get("value"), "unit": "flows/s",
            "n_gpus": 1, "higher_is_better": True, "dtype": "fp16 autocast (reference default)", "data": "synthetic",
            "config": workload_config(args, clips=r.get("at_clips")), "ref_cuda": r}
    print(json.dumps(line))


# ------------------------------------------------------------------------------- our arm
def run_b200(args):
    import torch.distributed as dist
    from accflow_b200 import _lib
    from accflow_b200 import metrics
    from accflow_b200.networks import build_flow_estimator
    from accflow_b200.networks.AccFlow_ import AccFlow
    from accflow_b200.weights import make_state_dict

    torch.set_grad_enabled(False)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    kind = f"acc+{args.ofe}"
    model = AccFlow(build_flow_estimator(kind))
    model.load_state_dict(make_state_dict(kind, seed=2))
    model = model.to(dev).eval()
    model.iters = args.iters
    model.ofe.precision = args.precision
    model.warm_start = bool(args.warm_start)
    b = args.clips
    # clip-parallel sharding: the step's clips are dealt round-robin; with --total-clips each rank
    # sweeps its shard in micro-batches of b (the last one may be ragged)
    from accflow_b200.sharding import gather_clip_metrics, shard_clip_ids
    n_clips = args.total_clips if args.total_clips else world * b
    my_ids = shard_clip_ids(n_clips, rank, world)
    n_micro = -(-len(my_ids) // b)
    b = -(-len(my_ids) // n_micro)                      # even micro-steps (64 clips, --clips 9 -> 8 x 8, not 7 x 9 + 1)
    micro = [my_ids[i:i + b] for i in range(0, len(my_ids), b)]
    batches = [make_inputs(ids, args.size) for ids in micro]
    dev_sets = [([t.to(dev) for t in bt["imgs"]], bt["bflows"][-1].to(dev), bt["fflows"][-1].to(dev)) for bt in batches]
    host_sets = [[t.pin_memory() for t in bt["imgs"]] for bt in batches]
    host_out = [torch.empty(len(ids), 2, args.size, args.size).pin_memory() for ids in micro]
    host_epe = torch.empty(len(my_ids), 3).pin_memory()

    def step_resident():
        epes = []
        for imgs, bflow, fflow in dev_sets:
            flows = model(images=imgs, test_mode=False)
            epes.append(metrics.clip_epe(flows[-1], bflow, fflow))           # (b,3): fused occlusion mask + EPE kernel
        epe = epes[0] if len(epes) == 1 else torch.cat(epes)
        return gather_clip_metrics(epe, n_clips, rank, world)                 # the only collective: metric gather

    # End to end: every (micro-)step copies its frames from pinned host memory, reads its last flow and its EPE
    # triplets back.  The copy of the next micro-step runs on a copy stream while the current one computes (two
    # device input sets); the first copy of a timed region is exposed.
    copy_stream = torch.cuda.Stream(device=dev)
    bmax = max(len(ids) for ids in micro)
    in_sets = [[torch.empty(bmax, *t.shape[1:], device=dev) for t in host_sets[0]] for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "primed": False}

    def upload(k, j):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_used[k])            # the step that last read this set has consumed it
            for d, h in zip(in_sets[k], host_sets[j]):
                d[: h.shape[0]].copy_(h, non_blocking=True)
            ev_ready[k].record(copy_stream)

    def step_e2e(last=False):
        nm = len(micro)
        epes = []
        for j in range(nm):
            k = e2e_state["i"] & 1
            if not e2e_state["primed"]:
                upload(k, j)
                e2e_state["primed"] = True
            final = last and j == nm - 1
            if not final:
                upload(k ^ 1, (j + 1) % nm)               # next micro-step's frames, overlapped with this one's kernels
            cur = torch.cuda.current_stream()
            cur.wait_event(ev_ready[k])
            nb = len(micro[j])
            flows = model(images=[t[:nb] for t in in_sets[k]], test_mode=False)
            ev_used[k].record(cur)
            host_out[j].copy_(flows[-1], non_blocking=True)
            epes.append(metrics.clip_epe(flows[-1], dev_sets[j][1], dev_sets[j][2]))
            e2e_state["i"] += 1
            if final:
                e2e_state["primed"] = False
        host_epe.copy_(epes[0] if nm == 1 else torch.cat(epes), non_blocking=True)
        return flows

    def timed(fn, steps, warmup, sampler=None, mark_last=False):
        for w in range(warmup):
            fn(last=(w == warmup - 1)) if mark_last else fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = _lib.call("accflow_launch_count", 0)
        e0.record()
        for k in range(steps):
            fn(last=(k == steps - 1)) if mark_last else fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        launches = _lib.call("accflow_launch_count", 0) - launches0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, clocks

    sampler = ClockSampler(local) if rank == 0 else None
    # >= 3 warm-up steps always: calls 1-2 of a shape run eagerly and the third captures the CUDA graph (engine.GraphCache);
    # a capture inside the timed region would be a measurement of the capture
    args.warmup = max(args.warmup, 3)
    ms, launches, clocks = timed(step_resident, args.steps, args.warmup, sampler)
    flows_total = FLOWS_PER_CLIP * n_clips * args.steps
    value = flows_total / (ms / 1e3)
    ms_e2e, _, _ = timed(step_e2e, args.steps, 3 if len(micro) > 1 or args.total_clips else 1, mark_last=True)
    e2e_value = flows_total / (ms_e2e / 1e3)
    h2d = sum(t.numel() * 4 for hs in host_sets for t in hs)
    d2h = sum(t.numel() * 4 for t in host_out) + host_epe.numel() * 4

    # the benched path's own output (CUDA-graph replay, b clips, seed-2 weights) for the parity leg below
    bench_flows = model(images=dev_sets[0][0], test_mode=False)
    bench_epe = metrics.clip_epe(bench_flows[-1], dev_sets[0][1], dev_sets[0][2])
    torch.cuda.synchronize()
    ours_clip0 = [f[0:1].float().cpu() for f in bench_flows]
    ours_epe0 = bench_epe[0].float().cpu()

    # ---- roofline of the dominant kernel (implicit-GEMM convolution), measured live ----------
    eng = model.engine(dev)
    model.ofe.use_cuda_graph = False           # per-launch events need the eager path
    prof = eng.k.profile = []
    pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    pe0.record()
    step_resident()
    pe1.record()
    torch.cuda.synchronize()
    prof_step_ms = pe0.elapsed_time(pe1)
    eng.k.profile = None
    model.ofe.use_cuda_graph = True
    conv_ms = sum(a.elapsed_time(z) for a, z, _ in prof)
    conv_flop = sum(f for _, _, f in prof)
    pk, pk_kind = peaks()
    achieved = conv_flop / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    peak = pk["bf16_tflops_sustained"]
    kname = {"fp32": "conv_f32_kernel (implicit-GEMM conv, exact-fp32 FFMA path)",
             "bf16x3": "conv_tc_kernel (tcgen05 implicit-GEMM conv, bf16x3 split: 6 MMAs per algorithmic MAC)",
             "fp16x2": "conv_tc_kernel (tcgen05 implicit-GEMM conv, fp16x2 split: 3 MMAs per algorithmic MAC)",
             "fp16": "conv_tc_kernel (tcgen05 implicit-GEMM conv, fp16 products: the reference's autocast class)",
             "bf16": "conv_tc_kernel (tcgen05 implicit-GEMM conv, bf16 products)"}[args.precision]
    issued = achieved * {"bf16x3": 6, "fp16x2": 3}.get(args.precision, 1)
    traffic, traffic_note = None, None
    for name in (f"r4p_conv_tc_zr_{args.precision}_ncu_full.jsonl", f"r4h_conv_tc_zr_{args.precision}_ncu_full.jsonl",
                 f"r3w_conv_tc_zr_{args.precision}_ncu_full.jsonl", f"r3c_conv_tc_zr_{args.precision}_ncu_full.jsonl",
                 f"r2_conv_tc_zr_{args.precision}_ncu_full.jsonl",
                 f"r1c_conv_tc_zr_{args.precision}_ncu_full.jsonl",
                 f"r1b_conv_tc_zr_{args.precision}_ncu_full.jsonl"):
        tfile = os.path.join(ROOT, "profiles", name)
        if os.path.exists(tfile):
            t = json.loads(open(tfile).readline())
            traffic = (t["dram_read_MB"] + t["dram_write_MB"]) * 1e6
            traffic_note = (f"STATIC, not measured in this run: dram__bytes_read+write of one GRU z|r conv launch (1x5, "
                            f"{'18' if name.startswith(('r2_', 'r3c_', 'r3w_', 'r4h_', 'r4p_')) else '8'} pairs x 64x64; `ncu --set full`) from profiles/{name}")
            break
    roofline = {"bound": "tensor", "kernel": kname, "issued_mma_tflops": issued, "issued_frac": issued / pk["bf16_tflops_sustained"],
                "traffic_note": traffic_note,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": f"bf16 dense sustained, {pk_kind}", "launches": len(prof), "ms_in_step": conv_ms,
                "share_of_step": conv_ms / prof_step_ms,
                "share_note": "conv launches / whole step, both CUDA-event timed in one eager (non-graph) step",
                "flop_note": "FLOPs of the convolutions/GEMMs actually launched (2*MAC); the GRU's constant `inp` term is "
                             "evaluated once per frame, not once per iteration",
                "traffic": traffic}

    # ---- second metric of BASELINE.json: ms per GRU iteration (lookup + update block), graph-replayed ----
    def gru_iter_ms(pairs):
        from accflow_b200.engine import View
        ofe = eng.ofe
        hh = args.size // 8
        g = torch.Generator(device="cpu").manual_seed(7)
        mk = lambda c, f: View(f(torch.randn(pairs, hh, hh, c, generator=g)).to(dev).contiguous())
        st = ofe.prepare(mk(256, lambda t: t), mk(256, lambda t: t), mk(128, torch.tanh), mk(128, torch.relu),
                         args.size, args.size, f"gi{pairs}")
        took = {}
        for iters in (4, 16):
            run = lambda: ofe.iterate(st, iters, None, f"gi{pairs}")
            run(); run()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                run()
            graph.replay()
            a, z = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                graph.replay()
            z.record()
            torch.cuda.synchronize()
            took[iters] = a.elapsed_time(z) / 5
        return (took[16] - took[4]) / 12.0

    ms_iter = None
    if rank == 0 and args.size <= 512:
        ms_iter = {"pairs_3 (one clip, first accumulation step)": gru_iter_ms(3),
                   f"pairs_{3 * b} (this bench's batch)": gru_iter_ms(3 * b)}

    line = None
    ok = True
    if rank == 0:
        line = {"metric": "long-range flow pairs/sec", "value": value, "unit": "flows/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.total_clips else "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32"}.get(args.precision, args.precision), "data": "synthetic",
                "config": workload_config(args, b), "clips_per_s": value / FLOWS_PER_CLIP,
                "pair_evals_per_s": value / FLOWS_PER_CLIP * 11, "ms_per_gru_iter": ms_iter, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "flows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "note": "frames H2D from pinned memory, last flow + per-clip EPE triplets D2H, metric kernel included"},
                "gpu_launches": int(launches), "roofline": roofline,
                "hbm_peak_bytes": int(torch.cuda.max_memory_allocated(dev)),
                "necessary_tflops": NECESSARY_GFLOP_PER_CLIP_512 * (args.size / 512) ** 2 * (args.iters / 12)
                                    * n_clips * args.steps / ms}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"], line["parity"] = cpu_leg(args, ours_clip0, ours_epe0, batches[0])
            fp32_class = args.precision in ("fp32", "bf16x3", "fp16x2")
            par = line["parity"]
            if fp32_class:
                ok = par["max_abs_px"] < FLOW_TOL_PX and par["epe_delta_px"] < EPE_TOL_PX
                if not ok and par["epe_delta_px"] < EPE_TOL_PX:
                    # The accumulation has a hard threshold (getOcc's binary mask, AccFlow_.py:130-134): on some clips a
                    # 1e-6 input perturbation flips it for a few pixels and moves the REFERENCE's own output by > 1e-3 px
                    # locally.  Measure that on this clip before calling a difference of the same size a failure.
                    par["oracle_self_sensitivity_px"] = oracle_self_sensitivity(args, batches[0])
                    par["note"] = ("max_abs_px is judged against 2x the oracle's own response to a 1e-6 input perturbation "
                                   "on this clip when that exceeds the 1e-3 px bar (ill-conditioned clip)")
                    ok = par["max_abs_px"] < max(FLOW_TOL_PX, 2.0 * par["oracle_self_sensitivity_px"])
            par["pass"] = bool(ok) if fp32_class else None
        if world == 1 and not args.no_ref_cuda:
            del dev_sets, in_sets
            torch.cuda.empty_cache()
            rc = ref_cuda_proxy(args, dev, ours_clip0=ours_clip0)
            line["ref_cuda"] = rc
            if rc.get("value"):
                line["ref_cuda"]["speedup_resident"] = value / rc["value"]
                line["ref_cuda"]["speedup_e2e"] = e2e_value / rc["value"]
                line["ref_cuda"]["north_star_10x"] = "met" if e2e_value / rc["value"] >= 10 else "missed"
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if not ok:
        sys.exit("parity failure: the benched path's flows differ from the oracle beyond the fp32-class bar")


def oracle_self_sensitivity(args, batch0):
    """max-abs change of the CPU oracle's flows for clip 0 under a 1e-6 input perturbation (conditioning of the clip)."""
    from accflow_b200.weights import make_state_dict
    from oracle import flow_oracle as fo
    sd = make_state_dict(f"acc+{args.ofe}", seed=2)
    imgs = [t[0:1] for t in batch0["imgs"]]
    g = torch.Generator().manual_seed(0)
    pert = [im + 1e-6 * torch.randn(im.shape, generator=g) for im in imgs]
    kw = dict(warm_start=True) if args.warm_start else {}
    a = fo.accflow_forward(sd, imgs, args.iters, **kw)
    b = fo.accflow_forward(sd, pert, args.iters, **kw)
    return max(float((x - y).abs().max()) for x, y in zip(a, b))


def cpu_leg(args, ours_clip0, ours_epe0, batch0):
    """Oracle port on the host cores of the GPU box over ONE whole clip (the bounded sample of the workload): its wall time
    is the cpu_baseline, its flows are what the benched GPU path (graph replay, full batch) is checked against."""
    from accflow_b200.weights import make_state_dict
    from oracle import flow_oracle as fo
    from oracle import ops
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(f"acc+{args.ofe}", seed=2)
    imgs = [t[0:1] for t in batch0["imgs"]]
    t0 = time.time()
    if args.warm_start:
        ref = fo.accflow_forward(sd, imgs, args.iters, warm_start=True)
    else:
        ref = fo.accflow_forward(sd, imgs, args.iters)
    t = time.time() - t0
    bflow, fflow = batch0["bflows"][-1][0:1], batch0["fflows"][-1][0:1]
    occ, _ = ops.calc_occ_mask(bflow, fflow)
    e_ref = torch.stack(ops.cal_epe(ref[-1], bflow, occ)).reshape(-1)
    e_ours_cpu = torch.stack(ops.cal_epe(ours_clip0[-1], bflow, occ)).reshape(-1)
    parity = {"max_abs_px": max(float((a - b).abs().max()) for a, b in zip(ours_clip0, ref)),
              "epe_delta_px": float((e_ours_cpu - e_ref).abs().max()),
              "epe_kernel_vs_oracle_px": float((ours_epe0 - e_ref).abs().max()),
              "clips_checked": 1, "flows_checked": len(ref), "flow_max_px": float(max(r.abs().max() for r in ref)),
              "what": "clip 0 of the benched batch: CUDA-graph-replayed flows F(2->0)..F(6->0) vs the CPU oracle (fp32)",
              "tolerance": {"max_abs_px": FLOW_TOL_PX, "epe_delta_px": EPE_TOL_PX}}
    base = {"value": FLOWS_PER_CLIP / t, "unit": "flows/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"one whole {FRAMES}-frame {args.size}x{args.size} clip (11 pair-evals, 5 accumulation steps), {t:.1f}s"}
    return base, parity


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "reference-cuda":
        run_reference_cuda(a)
    else:
        run_b200(a)
