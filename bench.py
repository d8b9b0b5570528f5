#!/usr/bin/env python
"""Benchmark of the AccFlow hot path (BASELINE.json: long-range flow pairs/sec, 7-frame
512x512 clips, 12 GRU iterations per pair).

  python bench.py --gpus 1 --steps 10 --warmup 3            # this implementation
  python bench.py --impl reference --steps 1 --warmup 0     # reference algorithm on host CPU
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of AccFlow+RAFT backward accumulation over `--clips` synthetic CVO-shaped
clips per GPU (5 long-range flows per 7-frame clip).  Prints ONE JSON line on rank 0.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=9,
                    help="clips per GPU per step (9 clips = 18/27 pairs fill the 148 SMs with 7.8/11.7 tile rounds)")
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--iters", type=int, default=12)
    ap.add_argument("--ofe", default="raft", choices=["raft", "gma"])
    ap.add_argument("--precision", default=os.environ.get("ACCFLOW_PRECISION", "fp16x2"), choices=["fp32", "bf16x3", "fp16x2", "bf16"],
                    help="conv/GEMM arithmetic: fp16x2 / bf16x3 = tcgen05 split products (fp32-class, parity-gated at 1e-3 px)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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


# ------------------------------------------------------------------------------- reference arm
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


def workload_config(args, clips):
    return {"workload": f"AccFlow+{args.ofe.upper()} backward accumulation, {FRAMES}-frame {args.size}x{args.size} CVO-shaped synthetic clip, "
                        f"{args.iters} iters/pair (BASELINE configs[1])",
            "clips_per_gpu_per_step": clips, "frames": FRAMES, "size": args.size, "iters": args.iters,
            "weights": "seeded random (accflow_b200.weights, seed 2)",
            "l2_policy": "working set per step (correlation volumes + activations) exceeds the 126 MB L2"}


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
    b = args.clips
    # clip-parallel sharding: the step's world*b clips are dealt round-robin
    from accflow_b200.sharding import gather_clip_metrics, shard_clip_ids
    n_clips = world * b
    batch = make_inputs(shard_clip_ids(n_clips, rank, world), args.size)
    host_imgs = [t.pin_memory() for t in batch["imgs"]]
    host_out = torch.empty(b, 2, args.size, args.size).pin_memory()
    dev_imgs = [t.to(dev) for t in batch["imgs"]]
    bflow, fflow = batch["bflows"][-1].to(dev), batch["fflows"][-1].to(dev)

    def step_resident():
        flows = model(images=dev_imgs, test_mode=False)
        epe = metrics.clip_epe(flows[-1], bflow, fflow)                       # (b,3): fused occlusion mask + EPE kernel
        return gather_clip_metrics(epe, n_clips, rank, world)                 # the only collective: metric gather

    # End to end: every step copies its frames from pinned host memory and reads its last flow back.  The copy of
    # step i+1 runs on a copy stream while step i computes (two device input sets); the first step's copy is exposed.
    copy_stream = torch.cuda.Stream(device=dev)
    in_sets = [[torch.empty_like(t, device=dev) for t in host_imgs] for _ in range(2)]
    ev_ready = [torch.cuda.Event() for _ in range(2)]
    ev_used = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"i": 0, "primed": False}

    def upload(k):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_used[k])            # the step that last read this set has consumed it
            for d, h in zip(in_sets[k], host_imgs):
                d.copy_(h, non_blocking=True)
            ev_ready[k].record(copy_stream)

    def step_e2e(last=False):
        k = e2e_state["i"] & 1
        if not e2e_state["primed"]:
            upload(k)
            e2e_state["primed"] = True
        if not last:
            upload(k ^ 1)                                  # next step's frames, overlapped with this step's kernels
        cur = torch.cuda.current_stream()
        cur.wait_event(ev_ready[k])
        flows = model(images=in_sets[k], test_mode=False)
        ev_used[k].record(cur)
        host_out.copy_(flows[-1], non_blocking=True)
        e2e_state["i"] += 1
        if last:
            e2e_state["primed"] = False
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
    ms, launches, clocks = timed(step_resident, args.steps, args.warmup, sampler)
    flows_total = FLOWS_PER_CLIP * b * world * args.steps
    value = flows_total / (ms / 1e3)
    ms_e2e, _, _ = timed(step_e2e, args.steps, 1, mark_last=True)
    e2e_value = flows_total / (ms_e2e / 1e3)
    h2d = sum(t.numel() * 4 for t in host_imgs)
    d2h = host_out.numel() * 4

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
             "bf16": "conv_tc_kernel (tcgen05 implicit-GEMM conv, bf16 products)"}[args.precision]
    issued = achieved * {"bf16x3": 6, "fp16x2": 3}.get(args.precision, 1)
    traffic, traffic_note = None, None
    tnew = os.path.join(ROOT, "profiles", f"r1c_conv_tc_zr_{args.precision}_ncu_full.jsonl")
    if not os.path.exists(tnew):
        tnew = os.path.join(ROOT, "profiles", f"r1b_conv_tc_zr_{args.precision}_ncu_full.jsonl")
    tfile = os.path.join(ROOT, "profiles", f"r1_conv_tc_zr_{args.precision}_ncu_full.json")
    note = ("dram__bytes_read+write of one GRU z|r conv launch (1x5, 384->256, 8 pairs x 64x64; `ncu --set full`) from "
            "profiles/{}; algorithmic bytes of that launch ~104 MB (operand planes 50 MB + weights 4 MB + h 17 MB "
            "read, z 17 MB + r*h planes 17 MB written), part of it served by the 126 MB L2")
    if os.path.exists(tnew):
        t = json.loads(open(tnew).readline())
        traffic = (t["dram_read_MB"] + t["dram_write_MB"]) * 1e6
        traffic_note = note.format(os.path.basename(tnew))
    elif os.path.exists(tfile):
        t = json.load(open(tfile))
        scale_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        traffic = sum(float(t[k]["value"]) * scale_b[t[k]["unit"]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        traffic_note = note.format(os.path.basename(tfile))
    roofline = {"bound": "tensor", "kernel": kname, "issued_mma_tflops": issued, "issued_frac": issued / pk["bf16_tflops_sustained"],
                "traffic_note": traffic_note,
                "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                "peak_kind": f"bf16 dense sustained, {pk_kind}", "launches": len(prof), "ms_in_step": conv_ms,
                "share_of_step": conv_ms / prof_step_ms,
                "share_note": "conv launches / whole step, both CUDA-event timed in one eager (non-graph) step",
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

    ms_iter = {"pairs_3 (one clip, first accumulation step)": gru_iter_ms(3),
               f"pairs_{3 * b} (this bench's batch)": gru_iter_ms(3 * b)} if rank == 0 else None

    line = None
    if rank == 0:
        line = {"metric": "long-range flow pairs/sec", "value": value, "unit": "flows/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None,
                "dtype": {"fp32": "f32", "bf16x3": "bf16x3", "fp16x2": "fp16x2", "bf16": "bf16"}[args.precision], "data": "synthetic",
                "config": workload_config(args, b), "clips_per_s": value / FLOWS_PER_CLIP,
                "pair_evals_per_s": value / FLOWS_PER_CLIP * 11, "ms_per_gru_iter": ms_iter, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "flows/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": int(launches), "roofline": roofline,
                "necessary_tflops": NECESSARY_GFLOP_PER_CLIP_512 * (args.size / 512) ** 2 * b * world * args.steps / ms}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args)
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_baseline(args):
    """Oracle port timed on the host cores of the GPU box, on a bounded sample of the workload."""
    from accflow_b200.weights import make_state_dict
    from oracle import flow_oracle as fo
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = make_state_dict(f"acc+{args.ofe}", seed=2)
    imgs = make_inputs([0], args.size)["imgs"][:3]          # first accumulation step: 3 pair-evals, 1 flow
    t0 = time.time()
    fo.accflow_forward(sd, imgs, args.iters)
    t = time.time() - t0
    # a full 7-frame clip is 11 pair-evals + 5 accumulation steps; this sample is 3 + 1
    est_clip = t * (11.0 / 3.0)
    return {"value": FLOWS_PER_CLIP / est_clip, "unit": "flows/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"first accumulation step of one {args.size}x{args.size} clip (3 of 11 pair-evals, {t:.1f}s), scaled x11/3 to a clip"}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
