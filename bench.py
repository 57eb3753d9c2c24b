#!/usr/bin/env python
"""Headline benchmark: ViT-B/16 train images/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dino_vitb16] [--batch 128]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's torch CPU path (oracle port) on the host cores

A step is one fine-tune iteration of the reference loop (utils_network.py:406-452): forward, CrossEntropyLoss,
zero_grad / backward, SGD(momentum 0.9, lr 1e-3) step, on synthetic 224x224 images with random-init weights.
`value` is whole-job images/sec with the batch already resident in HBM; `e2e` is the same loop fed from pinned host
memory with the loss read back every step. One JSON line is printed by rank 0.
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

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (constructor, per-GPU batch, image size, tokens N, D, depth, heads, patch)
    "dino_vitb16": ("dino_vitb16", 128, 224, 197, 768, 12, 12, 16),
    "dino_vitb8": ("dino_vitb8", 64, 224, 785, 768, 12, 12, 8),
    "dino_vits16": ("dino_vits16", 128, 224, 197, 384, 12, 6, 16),
}


def fwd_flops_per_image(N, D, L, P, C=3):
    n = N - 1
    return 2 * n * C * P * P * D + L * (2 * N * D * 3 * D + 4 * N * N * D + 2 * N * D * D + 16 * N * D * D)


def gemm_flops_per_image(N, D, L, P, C=3):
    """Algorithmic FLOPs of the tcgen05 GEMM launches per image per train step: Linear fwd + dgrad + wgrad (3x) for
    qkv/proj/fc1/fc2, PatchEmbed fwd + wgrad (2x; images need no grad)."""
    n = N - 1
    lin = L * (2 * N * D * 3 * D + 2 * N * D * D + 16 * N * D * D)
    return 3 * lin + 2 * (2 * n * C * P * P * D)


def gemm_bytes_per_image(N, D, L, P, C=3):
    """Algorithmic DRAM bytes of the GEMM launches per image per train step (bf16 operands and saved activations, fp32
    residual stream and weight gradients; weights counted once per launch, amortised over the batch elsewhere)."""
    a = N * D  # elements of one [N, D] activation
    fwd = (2 * a + 6 * a) + (2 * a + 4 * a + 4 * a) + (2 * a + 8 * a + 8 * a) + (8 * a + 4 * a + 4 * a)
    dgrad = (2 * a + 8 * a + 8 * a) + (8 * a + 2 * a) + (2 * a + 2 * a) + (6 * a + 2 * a)
    wgrad = (2 * a + 8 * a) + (8 * a + 2 * a) + (2 * a + 2 * a) + (6 * a + 2 * a)
    return L * (fwd + dgrad + wgrad)


def measured_gemm_traffic():
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the GEMM launches of one step, from the
    committed ncu capture of this same command (profiles/r01_gemm_traffic.json, scripts/gpu_profile.sh)."""
    path = os.path.join(ROOT, "profiles", "r01_gemm_traffic.json")
    try:
        with open(path) as f:
            return float(json.load(f)["traffic_per_launch_bytes"])
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_burst=p["bf16_tflops"], bf16_sustained=p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                    hbm=p["hbm_gbs"], source="measured")
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


class ClockSampler:
    """Samples nvidia-smi SM clocks and throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "25"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=0.0, t1=float("inf")):
        """Summarise the samples taken in the wall-clock window [t0, t1] (the timed region)."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        window = [r for t, r in self.rows if t0 <= t <= t1] or [r for _, r in self.rows[-3:]]
        for r in window:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def synth_batch(B, size, seed, device):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn((B, 3, size, size), generator=g)
    y = torch.randint(0, 10, (B,), generator=g)
    return x.to(device), y.to(device)


def run_reference(args):
    """--impl reference: the reference's own (CPU, fp32, eager PyTorch) path = the oracle restatement, since the DINO
    code the reference pulls from torch.hub is not vendored. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    from oracle import train_step as ots
    name, _, size, N, D, L, H, P = WORKLOADS[args.workload]
    bs = args.ref_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, opt = ots.build(name, seed=0)
    x, y = synth_batch(bs, size, 0, "cpu")
    for _ in range(args.warmup):
        ots.step(model, opt, x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = ots.step(model, opt, x, y)
    dt = time.perf_counter() - t0
    ips = bs * args.steps / dt
    sample = f"{name} fine-tune step (fwd+CE+bwd+SGD) fp32 torch CPU, batch {bs} per step (bounded sample of the bs-128 workload)"
    line = {
        "impl": "reference", "metric": "ViT-B/16 train images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{name} fine-tune 224x224 (reference CPU path, oracle port)", "batch_per_step": bs},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": float(loss),
    }
    _emit(line)


def cpu_baseline(args):
    from oracle import train_step as ots
    name, _, size, *_ = WORKLOADS[args.workload]
    bs = args.ref_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, opt = ots.build(name, seed=0)
    x, y = synth_batch(bs, size, 0, "cpu")
    ots.step(model, opt, x, y)
    t0 = time.perf_counter()
    steps = 0
    while steps < 3 or (time.perf_counter() - t0 < 12 and steps < 40):
        ots.step(model, opt, x, y)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": bs * steps / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{name} fwd+CE+bwd+SGD fp32 torch CPU, batch {bs}, 1 warm-up + {steps} timed steps "
                      f"({dt:.1f} s) on the GPU box host"}


def _emit(line: dict) -> None:
    """Write THE one JSON line to the process's original stdout (see _claim_stdout)."""
    os.write(_STDOUT_FD, (json.dumps(line) + "\n").encode())


_STDOUT_FD = 1


def _claim_stdout() -> None:
    """Keep stdout clean for the single JSON line: libraries (NCCL prints its version banner on stdout) and stray
    prints are redirected to stderr; _emit() writes to the saved descriptor."""
    global _STDOUT_FD
    sys.stdout.flush()
    _STDOUT_FD = os.dup(1)
    os.dup2(2, 1)


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dino_vitb16", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sync-read", action="store_true",
                    help="e2e: read each step's loss with a blocking .item() right after enqueueing it (A/B; default: "
                         "the read of step i happens after step i+1 has been enqueued)")
    ap.add_argument("--overlap", action="store_true",
                    help="DP: per-block all-reduces overlapped with backward on NCCL_MAX_CTAS=--nccl-ctas thread blocks, "
                         "GEMM grids sized for the remaining SMs. Default: ONE all-reduce over the gradient arena after "
                         "backward on all SMs (measured faster on 2 B200: 19.7-19.8 vs 20.05-20.25 ms/step)")
    ap.add_argument("--no-overlap", action="store_true", help="(default behaviour; kept for older command lines)")
    ap.add_argument("--nccl-ctas", type=int, default=4,
                    help="DP: thread blocks left to the overlapped NCCL all-reduce (0 = NCCL default, GEMMs use every SM)")
    ap.add_argument("--no-graph", action="store_true", help="single GPU: launch every kernel eagerly (no CUDA graph)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch.distributed as dist
    from vit_torch_b200 import models, ops, train
    from vit_torch_b200.dist import GradAllReducer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        from vit_torch_b200.dist import configure_sm_partition
        if args.overlap:
            configure_sm_partition(args.nccl_ctas)
        dist.init_process_group("nccl", device_id=dev)

    name, bs_default, size, N, D, L, H, P = WORKLOADS[args.workload]
    bs = args.batch or bs_default
    torch.manual_seed(0)
    model = getattr(models, name)(pretrained=False).to(dev)
    train.reset_parameters_like_zoo(model)            # models/vision_all.py:157-158 (pretrained=False regime)
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    trainer = train.Trainer(model, lr=1e-3, momentum=0.9, reducer=GradAllReducer(model, overlap=args.overlap)
                            if world > 1 else None, graph=(not args.no_graph))

    x_dev, y_dev = synth_batch(bs, size, 1000 + rank, dev)
    x_host = torch.empty((bs, 3, size, size), dtype=torch.float32).pin_memory()
    y_host = torch.empty((bs,), dtype=torch.int64).pin_memory()
    x_host.copy_(x_dev.cpu())
    y_host.copy_(y_dev.cpu())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, finish=None):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            out = step_fn()
        if finish is not None:
            out = finish()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, out

    def step_resident():
        return trainer.step(x_dev, y_dev)

    # e2e input pipeline: every step's batch goes pinned host -> device inside the timed region, on a copy stream so
    # that the transfer of batch i+1 overlaps step i (what DataLoader(pin_memory=True) + non_blocking copies give the
    # reference loop, utils_network.py:409-410); the step waits on the copy's event. Every step's loss is read back
    # on the host inside the timed region: step i's value travels through a pinned buffer and is read right after
    # step i+1 has been enqueued (the last one before the closing event), so the GPU never idles behind the read.
    copy_stream = torch.cuda.Stream()
    stage_x = [torch.empty_like(x_dev) for _ in range(2)]
    stage_y = [torch.empty_like(y_dev) for _ in range(2)]
    stage_ev = [torch.cuda.Event() for _ in range(2)]
    e2e_state = {"it": 0, "primed": False, "pending": None, "losses": []}
    loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event() for _ in range(2)]

    def read_pending():
        k = e2e_state["pending"]
        if k is not None:
            loss_ev[k].synchronize()
            e2e_state["losses"].append(float(loss_host[k]))      # device -> host read of the step result
            e2e_state["pending"] = None
        return e2e_state["losses"][-1] if e2e_state["losses"] else None

    def prefetch(k):
        copy_stream.wait_stream(torch.cuda.current_stream())   # the buffer's previous consumer has been enqueued
        with torch.cuda.stream(copy_stream):
            stage_x[k].copy_(x_host, non_blocking=True)
            stage_y[k].copy_(y_host, non_blocking=True)
            stage_ev[k].record(copy_stream)

    def step_e2e():
        k = e2e_state["it"] & 1
        if not e2e_state["primed"]:
            prefetch(k)
            e2e_state["primed"] = True
        torch.cuda.current_stream().wait_event(stage_ev[k])
        st = trainer.static_inputs() if trainer.use_graph else None
        if st is not None:                                # captured step reads fixed buffers: device-to-device hand-over
            st[0].copy_(stage_x[k], non_blocking=True)
            st[1].copy_(stage_y[k], non_blocking=True)
            prefetch(k ^ 1)                               # next batch's H2D overlaps this step
            loss = trainer.step_static()
        else:
            prefetch(k ^ 1)
            loss = trainer.step(stage_x[k], stage_y[k])
        loss_host[k].copy_(loss, non_blocking=True)
        loss_ev[k].record()
        read_pending()                                    # the previous step's loss (this step is already enqueued)
        e2e_state["pending"] = k
        if args.e2e_sync_read:
            read_pending()
        e2e_state["it"] += 1
        return None

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~1 s to come up: start it before the warm-up
    for _ in range(max(args.warmup, 3)):                  # (graph mode: 2 eager steps, then the capture step)
        step_resident()
    l0 = ops.launch_count
    t_begin = time.time()
    ms, loss = timed(step_resident, args.steps)
    t_end = time.time()
    launches = (ops.launch_count - l0) + trainer.extra_launches_per_step * args.steps
    if trainer.use_graph and trainer.launches_per_step:
        launches = trainer.launches_per_step * args.steps  # replays do not pass through the Python launch counter
    value = world * bs * args.steps / (ms * 1e-3)

    e2e = None
    if not args.no_e2e:
        # stand-alone rate of the pinned host -> device copy on this box (reported next to e2e: when it is low, the
        # transfer of batch i+1 no longer hides completely behind step i)
        hs, he = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        hs.record()
        stage_x[0].copy_(x_host, non_blocking=True)
        he.record()
        torch.cuda.synchronize()
        h2d_gbs = x_host.numel() * 4 / (hs.elapsed_time(he) * 1e-3) / 1e9
        for _ in range(2):
            step_e2e()
        read_pending()
        torch.cuda.synchronize()
        e2e_state["primed"] = False                       # the first timed step issues its own H2D copy
        e2e_state["losses"] = []
        ms_e, _ = timed(step_e2e, args.steps, finish=read_pending)
        assert len(e2e_state["losses"]) == args.steps     # every step's loss reached the host inside the timed region
        e2e = {"value": world * bs * args.steps / (ms_e * 1e-3), "unit": "images/s",
               "h2d_bytes_per_step": x_host.numel() * 4 + y_host.numel() * 8, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e / args.steps, "last_loss": e2e_state["losses"][-1],
               "h2d_gbs_standalone": h2d_gbs}

    # roofline of the dominant kernel family (tcgen05 GEMM): CUDA events around every GEMM launch of the same loop
    pk = peaks()
    ops.gemm_timing_begin()
    barrier()
    for _ in range(args.steps):
        trainer._step_eager(x_dev, y_dev)                 # per-launch events need eager launches
    barrier()
    gemm_ms, gemm_launches = ops.gemm_timing_end()
    gemm_fl = gemm_flops_per_image(N, D, L, P) * bs * args.steps
    achieved = gemm_fl / (gemm_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel (tcgen05, all Linear fwd/dgrad/wgrad + PatchEmbed)",
                "achieved": achieved, "peak": pk["bf16_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_sustained"], "frac_of_burst_peak": achieved / pk["bf16_burst"],
                "peak_source": pk["source"] + " (sustained cuBLAS bf16, kernel timed inside a long step)",
                "traffic": measured_gemm_traffic() if (args.workload == "dino_vitb16" and bs == 128) else None,
                "traffic_unit": "bytes per launch (ncu dram read+write, profiles/r01_gemm_traffic.json)",
                "algorithmic_bytes_per_launch": gemm_bytes_per_image(N, D, L, P) * bs * args.steps / max(gemm_launches, 1),
                "launches": gemm_launches, "avg_launch_us": gemm_ms * 1e3 / max(gemm_launches, 1),
                "share_of_step": gemm_ms / (ms if ms > 0 else 1.0)}

    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    if rank == 0:
        step_fl = 3 * fwd_flops_per_image(N, D, L, P) * bs
        line = {
            "metric": "ViT-B/16 train images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{name} fine-tune (fwd+CE+bwd+SGD momentum 0.9) 224x224 synthetic, random init",
                       "batch_per_gpu": bs, "global_batch": bs * world, "tokens": N, "parallelism": f"dp{world}",
                       "l2": "working set per step (>8 GB of activations) far exceeds the 126 MB L2; no explicit flush",
                       "numerics": "bf16 GEMM/attention operands, fp32 accumulate, fp32 residual stream + master weights",
                       "launch": ("eager launches" if not trainer.use_graph else
                                  "whole step captured in one CUDA graph" if trainer.opt_in_graph else
                                  "forward + loss + backward captured in one CUDA graph; all-reduce and optimiser "
                                  "kernel launched after each replay"),
                       "grad_allreduce": ("none (1 GPU)" if world == 1 else
                                          f"per-block NCCL all-reduces overlapped with backward, NCCL_MAX_CTAS={args.nccl_ctas}"
                                          if args.overlap else
                                          "one NCCL all-reduce over the fp32 gradient arena after backward, all SMs")},
            "step_tflops_per_gpu": step_fl / (ms / args.steps * 1e-3) / 1e12,
            "step_frac_of_bf16_peak": step_fl / (ms / args.steps * 1e-3) / 1e12 / pk["bf16_sustained"],
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "loss": float(loss),
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
