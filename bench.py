#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: UNet-m64 SubmConv fwd+bwd voxels/sec (synthetic ScanNet-shaped scenes).

    python bench.py --gpus 1 --steps 5 --warmup 3            # this repo's CUDA path (N=1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W               # scene-sharded data parallel, NCCL grad all-reduce
    python bench.py --impl reference ...                     # the reference's own CPU arithmetic (oracle/_ref)

A "step" is one training step of the OccuSeg sparse backbone (InputLayer -> SubmConv 3->64 -> scn.UNet
[64..384], reps 1, residual -> BatchNormReLU -> OutputLayer; examples/ScanNet/model.py:657-691) on a batch of
8 synthetic 2 cm scenes of ~250k voxels each (BASELINE.json configs[2]): rulebook construction, forward,
backward (dgrad + wgrad), gradient all-reduce when N>1, Adam update.  value = level-0 active voxels of all
ranks / step time.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--preset", default="S250k")
    ap.add_argument("--scenes", type=int, default=8, help="scenes per GPU per step")
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "tf32", "fp32"])
    ap.add_argument("--config", type=int, default=3, choices=[2, 3, 5],
                    help="BASELINE.json configs: 3 = UNet m=64 training step, 8 x S250k (the metric's configuration, default); "
                         "2 = UNet m=32 inference on one S250k scene; 5 = UNet m=64 fwd+bwd on one 1M-voxel scene")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-precision-sweep", action="store_true", help="skip the tf32 / fp32 side measurements")
    a = ap.parse_args()
    if a.config == 2:
        a.preset, a.scenes, a.m = "S250k", 1, 32
    elif a.config == 5:
        a.preset, a.scenes, a.m = "S1M", 1, 64
    return a


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "measured"
    except Exception:
        return 6650.0, 1590.0, "fallback"


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.FIELDS}",
                                       "--format=csv,noheader,nounits", "-lms", "100"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if sm:
            sm.sort()
            out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


# =========================================================================================== reference arm
def cpu_workload(preset, m, max_seconds=40.0):
    """The reference's CPU arithmetic on a bounded sample of the workload: ONE scene instead of eight
    (a smaller preset if a pass would exceed max_seconds)."""
    import numpy as np
    from occuseg_b200 import scenes
    from oracle.unet_workload import Workload
    cores = os.cpu_count() or 1
    for ps in (preset, "S100k", "small"):
        coords, _ = scenes.make_batch(ps, (0,))
        w = Workload(coords, 1, m=m, levels=6)
        w.set_threads(cores)
        t = w.step()                       # warm-up pass doubles as the size probe
        if t <= max_seconds or ps == "small":
            return w, ps, cores
    return w, ps, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, ps, cores = cpu_workload(args.preset, args.m)
    for _ in range(max(args.warmup - 1, 0)):
        w.step()
    times = [w.step() for _ in range(args.steps)]
    total = sum(times)
    val = w.voxels * args.steps / total
    sample = f"1 scene ({ps}, {w.voxels} voxels) of the {args.scenes}-scene step: every UNet-m{args.m} sparse layer fwd+bwd"
    line = {
        "impl": "reference", "metric": "UNet-m64 SubmConv fwd+bwd voxels/sec", "value": val, "unit": "voxels/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"OccuSeg UNet m={args.m} fwd+bwd, {args.scenes} x {args.preset} scenes per GPU",
                   "preset": args.preset, "scenes_per_gpu": args.scenes, "m": args.m},
        "cpu_baseline": {"value": val, "unit": "voxels/s", "cores": cores, "kind": w.kind, "sample": sample},
        "e2e": {"value": val, "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# =========================================================================================== this repo's arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import occuseg_b200.sparseconvnet as scn
    from occuseg_b200 import _lib, scenes
    from occuseg_b200.backbone import SparseBackbone
    from occuseg_b200.ddp import BucketedGradAllReduce, FlatGradAllReduce

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("SCN_BENCH_BACKEND") == "gloo":      # debugging aid: isolate NCCL's effect on the device-timed step
            dist.init_process_group("gloo")
        elif os.environ.get("SCN_BENCH_LAZY_NCCL"):
            dist.init_process_group("nccl")
        else:
            dist.init_process_group("nccl", device_id=dev)
    scn.set_precision(args.precision)
    if os.environ.get("SCN_BENCH_VERBOSE"):
        sys.stderr.write(f"[rank {rank}] cpu affinity after init: {len(os.sched_getaffinity(0))} cpus {sorted(os.sched_getaffinity(0))[:8]}...\n")

    torch.manual_seed(1234)                       # identical initial weights on every rank
    net = SparseBackbone(m=args.m, levels=6).to(dev)
    inference = args.config == 2
    if inference:
        net.eval()
    opt = None if inference else torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
    Reducer = FlatGradAllReduce if os.environ.get("SCN_DDP", "bucketed") == "flat" else BucketedGradAllReduce
    reducer = Reducer(net.parameters(), world) if (world > 1 and not inference and os.environ.get("SCN_DDP") != "none") else None

    # ---- synthetic batch: rank r gets seeds r*scenes .. r*scenes+scenes-1 (weak scaling)
    seed0 = int(os.environ.get("SCN_BENCH_SEED0", "0"))          # debugging aid: run another rank's scenes on one GPU
    seeds = tuple(seed0 + rank * args.scenes + i for i in range(args.scenes))
    coords_np, feats_np = scenes.make_batch(args.preset, seeds)
    coords_host = torch.from_numpy(coords_np).pin_memory()
    feats_host = torch.from_numpy(feats_np).pin_memory()
    coords_dev = coords_host.to(dev)
    feats_dev = feats_host.to(dev)
    B = args.scenes

    def step(coords, feats):
        if inference:
            with torch.no_grad():
                return net([coords, feats, None, B]).square().mean()
        out = net([coords, feats, None, B])
        loss = out.square().mean()
        loss.backward()
        if reducer is not None:
            reducer.finish()
        opt.step()
        # with a reducer the gradients are views into its flat buffer and are zeroed in place; a single GPU drops them, so
        # that the next backward pass hands its gradient tensors over instead of adding them to zeros (209 launches less)
        opt.zero_grad(set_to_none=reducer is None)
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    host_ms = []

    def timed(fn, steps):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        e1.record()
        host_ms.append((time.perf_counter() - t0) * 1e3 / steps)     # host time to ENQUEUE a step (no device wait inside)
        sync_all()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # ---- warm-up (also reveals the voxel count)
    for _ in range(args.warmup):
        step(coords_dev, feats_dev)
    torch.cuda.synchronize()
    with torch.no_grad():
        n_vox = net.input([coords_dev, feats_dev, None, B]).features.shape[0]
    vox = torch.tensor([n_vox], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(vox)
    total_voxels = float(vox.item())

    # ---- timed region 1: inputs resident in HBM (no per-launch instrumentation inside)
    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = _lib.launch_count()
    ms = timed(lambda: step(coords_dev, feats_dev), args.steps)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if sampler else None

    # ---- the same steps again with CUDA events around every library launch: per-kernel-family times for the
    # roofline line (the events cost a few ms per step, which is why this pass is not the one that is reported)
    _lib.profile(True)
    ms_prof = timed(lambda: step(coords_dev, feats_dev), args.steps)
    prof = _lib.profile_read()
    _lib.profile(False)

    if os.environ.get("SCN_BENCH_VERBOSE"):      # per-rank view (debugging aid; the JSON line is rank 0's)
        sys.stderr.write(f"[rank {rank}] ms/step {ms / args.steps:.2f} instrumented {ms_prof / args.steps:.2f} "
                         + " ".join(f"{k}={v['ms'] / args.steps:.2f}" for k, v in prof.items() if v["launches"]) + "\n")

    # ---- timed region 2: end to end from pinned host buffers, loss read back every step.  Every step's inputs are
    # copied host->device inside the timed region; the copy of step i+1 is issued on a copy stream while step i
    # computes (a double-buffered loader), so it is hidden behind the device work instead of preceding it.
    copy_stream = torch.cuda.Stream()
    pending = {}

    def issue_copy():
        with torch.cuda.stream(copy_stream):
            c = coords_host.to(dev, non_blocking=True)
            f = feats_host.to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        pending["next"] = (c, f, ev)

    def e2e_step():
        c, f, ev = pending.pop("next")
        torch.cuda.current_stream().wait_event(ev)
        c.record_stream(torch.cuda.current_stream())
        f.record_stream(torch.cuda.current_stream())
        issue_copy()                                   # next step's inputs, overlapped with this step
        return float(step(c, f).item())

    issue_copy()
    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    pending.clear()
    h2d = coords_host.numel() * 8 + feats_host.numel() * 4
    d2h = 4

    # ---- side measurement: the same step on the tf32 tiles and on the exact fp32 path (rel 1e-5), two steps each
    precision_ms = {args.precision: ms / args.steps}
    if world == 1 and not args.no_precision_sweep:
        for prec in ("tf32", "fp32"):
            if prec == args.precision:
                continue
            try:
                scn.set_precision(prec)
                step(coords_dev, feats_dev)
                precision_ms[prec] = timed(lambda: step(coords_dev, feats_dev), 2) / 2
            except Exception as e:      # never lose the headline number to a side measurement
                precision_ms[prec] = f"failed: {e}"
            finally:
                scn.set_precision(args.precision)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm_peak, tc_peak, peak_src = peaks()
    kinds = {k: v for k, v in prof.items() if v["launches"]}
    top = max(kinds, key=lambda k: kinds[k]["ms"]) if kinds else None
    roofline = None
    if top:
        t = kinds[top]
        per_launch_ms = t["ms"] / t["launches"]
        achieved = (t["bytes"] / t["launches"]) / (per_launch_ms * 1e-3) / 1e9
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(top)
        except Exception:
            pass
        # two predictions of the launch time: gather-model bytes at the measured HBM copy peak, and the launch's FLOPs at the
        # sustained tensor peak of the operand type; the LARGER predicted time is the binding roofline
        tc_peak_used = tc_peak if args.precision == "bf16" else tc_peak / 2
        t_hbm = t["bytes"] / t["launches"] / (hbm_peak * 1e9) * 1e3
        t_tc = t["flops"] / t["launches"] / (tc_peak_used * 1e12) * 1e3 if t["flops"] else 0.0
        bound = "tensor" if t_tc > t_hbm else "hbm"
        if bound == "tensor":
            ach_tc = t["flops"] / (t["ms"] * 1e-3) / 1e12
            head = {"bound": "tensor", "achieved": ach_tc, "peak": tc_peak_used, "unit": "TFLOP/s", "frac": ach_tc / tc_peak_used}
        else:
            head = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak}
        roofline = {**head, "kernel": top, "traffic": traffic, "peak_source": peak_src,
                    "frac_hbm_gather_model": t_hbm / per_launch_ms, "frac_tensor": t_tc / per_launch_ms,
                    "predicted_ms_hbm": t_hbm, "predicted_ms_tensor": t_tc,
                    "launches": t["launches"], "avg_launch_ms": per_launch_ms,
                    "algorithmic_bytes_per_launch": t["bytes"] / t["launches"],
                    "share_of_step": t["ms"] / ms_prof, "instrumented_ms_per_step": ms_prof / args.steps,
                    "tensor_tflops": t["flops"] / (t["ms"] * 1e-3) / 1e12 if t["flops"] else None,
                    "tensor_peak_tflops": tc_peak_used,
                    "by_kernel_ms_per_step": {k: v["ms"] / args.steps for k, v in kinds.items()},
                    # the same algorithmic-bytes / CUDA-event-time ratio for every kernel family of the library
                    "families": {k: {"achieved_gbs": v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] else None,
                                     "frac": v["bytes"] / (v["ms"] * 1e-3) / 1e9 / hbm_peak if v["ms"] else None,
                                     "launches_per_step": v["launches"] / args.steps,
                                     "tflops": v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["flops"] and v["ms"] else None}
                                 for k, v in kinds.items()}}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            w, ps, cores = cpu_workload(args.preset, args.m)
            t = min(w.step() for _ in range(1))
            cpu = {"value": w.voxels / t, "unit": "voxels/s", "cores": cores, "kind": w.kind,
                   "sample": f"1 scene ({ps}, {w.voxels} voxels) of the {args.scenes}-scene step, every UNet-m{args.m} "
                             f"sparse layer fwd+bwd through the reference CPU code, {t:.1f} s"}
        except Exception as e:  # the baseline must never take the GPU number down with it
            cpu = {"value": None, "unit": "voxels/s", "cores": os.cpu_count(), "kind": "unavailable", "sample": str(e)}

    line = {
        "metric": "UNet-m64 SubmConv fwd+bwd voxels/sec", "value": total_voxels * args.steps / (ms * 1e-3),
        "unit": "voxels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16": "bf16 tensor-core operands (fp32 storage, fp32 accumulate)",
                  "tf32": "tf32 (fp32 storage, fp32 accumulate)", "fp32": "f32"}[args.precision], "data": "synthetic",
        "config": {"workload": (f"BASELINE.json configs[1]: OccuSeg UNet m={args.m} inference (eval mode, no grad), "
                                f"{args.scenes} x {args.preset} scene, rulebook build included" if inference else
                                f"BASELINE.json configs[{2 if args.config == 3 else 4}]: OccuSeg UNet m={args.m} (reps 1, residual, 6 levels) "
                                f"fwd+bwd+Adam, {args.scenes} x {args.preset} scenes per GPU, rulebook build included"),
                   "baseline_config": args.config,
                   "preset": args.preset, "scenes_per_gpu": args.scenes, "m": args.m,
                   "voxels_per_step": total_voxels, "parallelism": f"scene-sharded dp{world}",
                   "l2": "inputs larger than L2 (level-0 activations 0.5 GB per tensor); no explicit flush"},
        "e2e": {"value": total_voxels * args.steps / (ms_e2e * 1e-3), "unit": "voxels/s",
                "ms_per_step": ms_e2e / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "precision_ms": precision_ms, "host_enqueue_ms_per_step": host_ms[0] if host_ms else None,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
