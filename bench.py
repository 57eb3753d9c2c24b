#!/usr/bin/env python
"""Headline benchmark: ViT-B/16 train images/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload dino_vitb16] [--batch 128] [--input u8|bf16|f32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      # the reference's torch CPU path (oracle port) on the host cores

Workloads (BASELINE.json configs): dino_vitb16 (config 2, default), dino_vitb8 (3), cait_S24_224 (4),
dino_vitb16_lineareval (5), dino_vits16 (1 is its CPU form: `--impl reference --workload dino_vits16`).
A train step is one iteration of the reference loop (utils_network.py:406-452): forward, CrossEntropyLoss, zero_grad /
backward, SGD(momentum 0.9, lr 1e-3) step, on synthetic 224x224 images with random-init weights; a lineareval step is
the frozen backbone forward under no_grad plus the training step of the fc head (main.py:184-201).
`value` is whole-job images/sec with the batch already resident in HBM; `e2e` is the same loop fed from pinned host
memory through vit_torch_b200.data.DevicePrefetcher with the loss read back every step. One JSON line is printed by
rank 0.
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

# STL-10 statistics of the reference's Normalize (utils_datasets.py:586-589): the device input pipeline applies them
NORM_MEAN = [0.44671062065972217, 0.43980983983523964, 0.40664644709967324]
NORM_STD = [0.2603409782662331, 0.25657727311344447, 0.27126738145225493]

WORKLOADS = {
    # name: constructor, kind, per-GPU batch, image size, tokens N, D, depth, heads, patch
    "dino_vitb16": dict(ctor="dino_vitb16", kind="train", bs=128, size=224, N=197, D=768, L=12, H=12, P=16),
    "dino_vitb8": dict(ctor="dino_vitb8", kind="train", bs=64, size=224, N=785, D=768, L=12, H=12, P=8),
    "dino_vits16": dict(ctor="dino_vits16", kind="train", bs=128, size=224, N=197, D=384, L=12, H=6, P=16),
    "cait_S24_224": dict(ctor="cait_S24_224", kind="train", bs=128, size=224, N=196, D=384, L=24, H=8, P=16, cait=True),
    "dino_vitb16_lineareval": dict(ctor="dino_vitb16", kind="lineareval", bs=512, size=224, N=197, D=768, L=12, H=12,
                                   P=16),
}
HEAD_UNITS = [256, 128, 32, 10]     # --fc 256 128 32 + 10 labels (BASELINE config 1 / main.py:196-201)


# ------------------------------------------------------------------------------------------------------------------
# algorithmic work
# ------------------------------------------------------------------------------------------------------------------
def linear_flops_per_image(w):
    N, D, L = w["N"], w["D"], w["L"]
    f = L * (2 * N * D * 3 * D + 2 * N * D * D + 16 * N * D * D)
    if w.get("cait"):      # two class-attention blocks: q on 1 row, k/v on N+1 rows, proj + Mlp on 1 row
        f += 2 * (2 * D * D + 2 * 2 * (N + 1) * D * D + 2 * D * D + 16 * D * D)
    return f


def attn_flops_per_image(w):
    N, D, L, H = w["N"], w["D"], w["L"], w["H"]
    f = L * 4 * N * N * D
    if w.get("cait"):
        f += L * 2 * 2 * N * N * H * H + 2 * 4 * (N + 1) * D       # talking-heads mixes; class attention (1 query row)
    return f


def patch_flops_per_image(w, C=3):
    n = w["N"] - (0 if w.get("cait") else 1)
    return 2 * n * C * w["P"] * w["P"] * w["D"]


def fwd_flops_per_image(w):
    return patch_flops_per_image(w) + linear_flops_per_image(w) + attn_flops_per_image(w)


def step_flops_per_image(w):
    return fwd_flops_per_image(w) * (3 if w["kind"] == "train" else 1)


def family_work_per_image(w):
    """Algorithmic work of one step per image: tensor FLOPs of the GEMM family (Linear fwd + dgrad + wgrad, PatchEmbed
    fwd + wgrad) and of the attention family (fwd + 2x bwd, no credit for recompute), HBM bytes of the LayerNorm family
    (fwd: read fp32 x, write bf16 y = 6 B/elem; bwd: read bf16 dy + fp32 x + fp32 dres, write fp32 dx + bf16 copy = 16)."""
    train = w["kind"] == "train"
    N, D, L = w["N"], w["D"], w["L"]
    gemm = (3 if train else 1) * linear_flops_per_image(w) + (2 if train else 1) * patch_flops_per_image(w)
    attn = (3 if train else 1) * attn_flops_per_image(w)
    ln_calls = 2 * L + 1 + (4 if w.get("cait") else 0)
    ln_bytes = ln_calls * N * D * (6 + (16 if train else 0))
    return {"gemm": gemm, "attention": attn, "layernorm": ln_bytes}


OP_FAMILY = {"gemm": "gemm", "gemm_batched": "attention", "patch_embed_fwd": "gemm", "patch_embed_wgrad": "gemm",
             "attn_fwd": "attention", "attn_bwd": "attention", "th_mix_fwd": "attention", "th_mix_bwd": "attention",
             "th_scores": "attention", "th_apply": "attention", "th_apply_t": "attention",
             "class_attn_fwd": "attention", "class_attn_bwd": "attention", "layernorm_fwd": "layernorm",
             "layernorm_bwd": "layernorm", "layernorm_fwd_rows": "layernorm", "layernorm_bwd_rows": "layernorm"}


def measured_gemm_traffic():
    """Per-launch DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of the GEMM launches of one step, from the
    committed ncu capture of this same command (profiles/r01_gemm_traffic.json, scripts/gpu_profile.sh)."""
    for name in ("r02_gemm_traffic.json", "r01_gemm_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                return float(json.load(f)["traffic_per_launch_bytes"]), name
        except Exception:
            continue
    return None, None


def gemm_bytes_per_image(w):
    """Algorithmic DRAM bytes of the Linear GEMM launches per image per train step (bf16 operands and saved activations,
    fp32 residual stream and weight gradients; weights amortised over the batch)."""
    a = w["N"] * w["D"]
    fwd = (2 * a + 6 * a) + (2 * a + 4 * a + 4 * a) + (2 * a + 8 * a + 8 * a) + (8 * a + 4 * a + 4 * a)
    dgrad = (2 * a + 8 * a + 8 * a) + (8 * a + 2 * a) + (2 * a + 2 * a) + (6 * a + 2 * a)
    wgrad = (2 * a + 8 * a) + (8 * a + 2 * a) + (2 * a + 2 * a) + (6 * a + 2 * a)
    return w["L"] * (fwd + dgrad + wgrad)


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


def synth_batch(B, size, seed, fmt="f32"):
    """Synthetic host batch: uint8 pixels (what a decoded image is), or their ToTensor + Normalize form (what the
    reference's loader yields, utils_datasets.py:573-580) in fp32 / bf16."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    x8 = torch.randint(0, 256, (B, 3, size, size), generator=g, dtype=torch.uint8)
    y = torch.randint(0, 10, (B,), generator=g)
    if fmt == "u8":
        return x8, y
    mean = torch.tensor(NORM_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(NORM_STD).view(1, 3, 1, 1)
    x = (x8.float() / 255 - mean) / std
    return (x.to(torch.bfloat16) if fmt == "bf16" else x), y


# ------------------------------------------------------------------------------------------------------------------
# the reference arm / CPU baseline: the oracle port of the reference's torch CPU path
# ------------------------------------------------------------------------------------------------------------------
def _oracle_step_fn(workload, device, seed=0, autocast=False):
    """(model-ish, step(x, y) -> loss) for the oracle restatement of a workload. Test / baseline infrastructure."""
    import torch.nn as nn
    import torch.nn.functional as F
    from oracle import train_step as ots
    w = WORKLOADS[workload]
    torch.manual_seed(seed)
    if w.get("cait"):
        from oracle import cait as ocait
        model = ocait.create(w["ctor"], num_classes=10).to(device)
        opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)
        backbone = None
    else:
        model, opt = ots.build(w["ctor"], seed=seed, device=device)
        backbone = None
        if w["kind"] == "lineareval":
            backbone = model
            layers, fin = [], w["D"]
            for i, u in enumerate(HEAD_UNITS):           # models/vision_all.py:299-320
                last = i == len(HEAD_UNITS) - 1
                layers.append(nn.Linear(fin, u, bias=not last))
                if not last:
                    layers.append(nn.GELU())
                fin = u
            model = nn.Sequential(*layers).to(device)
            opt = torch.optim.SGD(model.parameters(), lr=1e-3, momentum=0.9)

    def step(x, y):
        with torch.autocast(device_type=torch.device(device).type, dtype=torch.bfloat16, enabled=autocast):
            if backbone is not None:
                with torch.no_grad():
                    f = backbone(x)
                out = model(f.float())
            else:
                out = model(x)
            loss = F.cross_entropy(out.float(), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        return loss.detach()
    return step


def run_reference(args):
    """--impl reference: the reference's own (CPU, fp32, eager PyTorch) path = the oracle restatement, since the DINO
    code the reference pulls from torch.hub is not vendored. Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w = WORKLOADS[args.workload]
    bs = args.ref_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _oracle_step_fn(args.workload, "cpu")
    x, y = synth_batch(bs, w["size"], 0, "f32")
    for _ in range(args.warmup):
        step(x, y)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        loss = step(x, y)
    dt = time.perf_counter() - t0
    ips = bs * args.steps / dt
    what = "fine-tune step (fwd+CE+bwd+SGD)" if w["kind"] == "train" else "lineareval step (frozen fwd + head train)"
    sample = (f"{args.workload} {what} fp32 torch CPU, batch {bs} per step (bounded sample of the bs-{w['bs']} workload)")
    line = {
        "impl": "reference", "metric": "ViT-B/16 train images/sec", "value": ips, "unit": "images/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": {"workload": f"{args.workload} 224x224 (reference CPU path, oracle port)", "batch_per_step": bs},
        "cpu_baseline": {"value": ips, "unit": "images/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": ips, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "loss": float(loss),
    }
    _emit(line)


def cpu_baseline(args):
    w = WORKLOADS[args.workload]
    bs = args.ref_batch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step = _oracle_step_fn(args.workload, "cpu")
    x, y = synth_batch(bs, w["size"], 0, "f32")
    step(x, y)
    t0 = time.perf_counter()
    steps = 0
    while steps < 3 or (time.perf_counter() - t0 < 12 and steps < 40):
        step(x, y)
        steps += 1
    dt = time.perf_counter() - t0
    return {"value": bs * steps / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{args.workload} step, fp32 torch CPU, batch {bs}, 1 warm-up + {steps} timed steps "
                      f"({dt:.1f} s) on the GPU box host"}


def gpu_eager_baseline(args, dev, bs):
    """Stock PyTorch eager on the SAME B200: the oracle modules (= the reference's math in
    plain torch.nn) in fp32 with TF32 off, and under autocast(bf16); same batch, same step."""
    w = WORKLOADS[args.workload]
    out = {}
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    x, y = synth_batch(bs, w["size"], 0, "f32")
    x, y = x.to(dev), y.to(dev)
    for name, ac in (("fp32_tf32_off", False), ("autocast_bf16", True)):
        try:
            step = _oracle_step_fn(args.workload, dev, autocast=ac)
            for _ in range(2):
                step(x, y)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            n = 5
            s.record()
            for _ in range(n):
                loss = step(x, y)
            e.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e) / n
            out[name] = {"images_per_s": bs / (ms * 1e-3), "ms_per_step": ms, "steps": n, "loss": float(loss)}
            del step
        except Exception as exc:   # e.g. out of memory for a large fp32 configuration: report, do not fail the bench
            out[name] = {"error": f"{type(exc).__name__}: {exc}"[:200]}
        torch.cuda.empty_cache()
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    out["what"] = ("oracle torch.nn modules (the reference's math), eager, cuBLAS/cuDNN/ATen kernels, batch "
                   f"{bs}, fp32 inputs resident in HBM, 2 warm-up + 5 timed steps, CUDA events")
    return out


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


class LinearEval(torch.nn.Module):
    """main.py:184-201 + utils_network.py:413-418: frozen backbone under no_grad, the fc head is the trained model."""

    def __init__(self, backbone, head):
        super().__init__()
        self.backbone, self.head = backbone, head
        for p in backbone.parameters():
            p.requires_grad_(False)

    def forward(self, x):
        with torch.no_grad():
            f = self.backbone(x)
        return self.head(f)


def build_model(args, dev):
    from vit_torch_b200 import cait, models, train, zoo
    w = WORKLOADS[args.workload]
    torch.manual_seed(0)
    if w.get("cait"):     # models/vision_all.py:184-213: create_model(...) + head [.., labels]; no reset (App. C.2)
        model = getattr(cait, w["ctor"])(pretrained=False, num_classes=1000)
        model.head = zoo.get_classifier_head(w["D"], HEAD_UNITS)
        pe = model.patch_embed
    else:
        model = getattr(models, w["ctor"])(pretrained=False)
        train.reset_parameters_like_zoo(model)            # models/vision_all.py:157-158 (pretrained=False regime)
        pe = model.patch_embed
        if w["kind"] == "lineareval":
            model = LinearEval(model, zoo.get_classifier_head(w["D"], HEAD_UNITS))
    model = model.to(dev)
    if args.input == "u8":
        pe.set_input_normalization(NORM_MEAN, NORM_STD)
    return model


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="dino_vitb16", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch (default: the BASELINE config's)")
    ap.add_argument("--input", default="u8", choices=["u8", "bf16", "f32"],
                    help="form of a batch as the loader hands it over: u8 = raw pixels, ToTensor + Normalize run on the "
                         "GPU (device input pipeline, default); bf16 / f32 = normalised on the host as the reference does")
    ap.add_argument("--ref-batch", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-families", action="store_true", help="skip the eager per-family timing pass")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-sync-read", action="store_true",
                    help="e2e: read each step's loss with a blocking .item() right after enqueueing it (A/B; default: "
                         "the read of step i happens after step i+1 has been enqueued)")
    ap.add_argument("--dp", default="auto", choices=["auto", "deferred", "overlap", "bf16", "split", "registered"],
                    help="DP gradient exchange: deferred = one fp32 all-reduce over the gradient arena after backward; "
                         "bf16 = the same in bf16; split = backward captured as two graphs, the first half's all-reduce "
                         "runs on --nccl-ctas thread blocks under the second graph; overlap = per-block all-reduces "
                         "overlapped with an eagerly launched backward")
    ap.add_argument("--overlap", action="store_true", help="(= --dp overlap)")
    ap.add_argument("--no-overlap", action="store_true", help="(default behaviour; kept for older command lines)")
    ap.add_argument("--nccl-ctas", type=int, default=4,
                    help="DP: thread blocks left to the overlapped NCCL all-reduce (0 = NCCL default, GEMMs use every SM)")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel eagerly (no CUDA graph)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return
    if args.overlap:
        args.dp = "overlap"

    import torch.distributed as dist
    from vit_torch_b200 import data, ops, train
    from vit_torch_b200.dist import GradAllReducer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    from vit_torch_b200.dist import default_mode
    dp_mode = default_mode(world) if (args.dp == "auto" or world == 1) else args.dp
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        from vit_torch_b200.dist import configure_sm_partition
        if dp_mode == "overlap":
            configure_sm_partition(args.nccl_ctas)
        dist.init_process_group("nccl", device_id=dev)

    w = WORKLOADS[args.workload]
    bs = args.batch or w["bs"]
    model = build_model(args, dev)
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    reducer = None
    if world > 1:
        reducer = GradAllReducer(model, overlap=(dp_mode == "overlap"), compress=(dp_mode == "bf16"),
                                 split=(dp_mode == "split"), split_ctas=args.nccl_ctas,
                                 register_arena=(dp_mode == "registered"))
    trainer = train.Trainer(model, lr=1e-3, momentum=0.9, reducer=reducer, graph=(not args.no_graph), strict_graph=True)

    x_host, y_host = synth_batch(bs, w["size"], 1000 + rank, args.input)
    x_host2, y_host2 = synth_batch(bs, w["size"], 5000 + rank, args.input)
    x_dev, y_dev = x_host.to(dev), y_host.to(dev)

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

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # nvidia-smi needs ~1 s to come up: start it before the warm-up
    for _ in range(max(args.warmup, 3)):                  # (graph mode: 2 eager steps, then the capture step)
        step_resident()
    if not args.no_graph and not trainer.use_graph and dp_mode != "overlap":   # (overlap issues NCCL from inside backward: eager by design)
        raise SystemExit("bench.py: the step was not captured in a CUDA graph (pass --no-graph to time eager launches)")
    l0 = ops.launch_count
    t_begin = time.time()
    ms, loss = timed(step_resident, args.steps)
    t_end = time.time()
    launches = (ops.launch_count - l0) + trainer.extra_launches_per_step * args.steps
    if trainer.use_graph and trainer.launches_per_step:
        launches = trainer.launches_per_step * args.steps  # replays do not pass through the Python launch counter
    value = world * bs * args.steps / (ms * 1e-3)
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None

    # ---- e2e: every step's batch travels pinned host -> device inside the timed region (the package's prefetcher:
    # double-buffered, copy stream), every step's loss is read on the host (step i's value is read after step i+1 has
    # been enqueued, the last one before the closing event)
    e2e = None
    if not args.no_e2e:
        hx = [x_host.pin_memory(), x_host2.pin_memory()]
        hy = [y_host.pin_memory(), y_host2.pin_memory()]
        hs, he = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tmp = torch.empty_like(x_dev)
        torch.cuda.synchronize()
        hs.record()
        tmp.copy_(hx[0], non_blocking=True)
        he.record()
        torch.cuda.synchronize()
        h2d_gbs = hx[0].numel() * hx[0].element_size() / (hs.elapsed_time(he) * 1e-3) / 1e9
        del tmp
        loss_host = [torch.empty((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        state = {"losses": [], "pending": None}

        def read_pending():
            k = state["pending"]
            if k is not None:
                loss_ev[k].synchronize()
                state["losses"].append(float(loss_host[k]))          # device -> host read of the step result
                state["pending"] = None
            return state["losses"][-1] if state["losses"] else None

        def run_e2e(n):
            batches = ((hx[i & 1], hy[i & 1]) for i in range(n))
            for i, (xb, yb) in enumerate(data.DevicePrefetcher(batches, dev)):
                k = i & 1
                ls = trainer.step(xb, yb)
                loss_host[k].copy_(ls, non_blocking=True)
                loss_ev[k].record()
                read_pending()                                        # the previous step's loss
                state["pending"] = k
                if args.e2e_sync_read:
                    read_pending()
            return read_pending()

        run_e2e(2)
        torch.cuda.synchronize()
        state["losses"] = []
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        run_e2e(args.steps)
        e.record()
        barrier()
        ms_e = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms_e], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = t.item()
        assert len(state["losses"]) == args.steps          # every step's loss reached the host inside the timed region
        e2e = {"value": world * bs * args.steps / (ms_e * 1e-3), "unit": "images/s",
               "h2d_bytes_per_step": hx[0].numel() * hx[0].element_size() + hy[0].numel() * 8, "d2h_bytes_per_step": 4,
               "ms_per_step": ms_e / args.steps, "last_loss": state["losses"][-1], "h2d_gbs_standalone": h2d_gbs,
               "input": args.input, "loader": "vit_torch_b200.data.DevicePrefetcher (pinned, 2 device slots, copy stream)"}

    # ---- per-family rooflines: CUDA events around every kernel-wrapper call of an EAGER pass of the same step (a graph
    # replay cannot be instrumented per kernel; the eager step runs the same kernels in the same order)
    pk = peaks()
    roofline, families = None, None
    if not args.no_families:
        eager = train.Trainer(model, lr=1e-3, momentum=0.9, reducer=None, graph=False) if world == 1 else None
        n_f = min(args.steps, 3)
        if eager is not None:
            eager.opt = trainer.opt          # same optimiser state
            eager.arena = trainer.arena
            for _ in range(1):
                eager._step_eager(x_dev, y_dev)
            torch.cuda.synchronize()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ops.op_timing_begin()
            s.record()
            for _ in range(n_f):
                eager._step_eager(x_dev, y_dev)
            e.record()
            per_op = ops.op_timing_end()
            eager_ms = s.elapsed_time(e) / n_f
            fam_ms, fam_n = {}, {}
            for name, (t, n) in per_op.items():
                fam = OP_FAMILY.get(name, "other")
                fam_ms[fam] = fam_ms.get(fam, 0.0) + t / n_f
                fam_n[fam] = fam_n.get(fam, 0) + n // n_f
            work = family_work_per_image(w)
            families = []
            for fam, bound in (("gemm", "tensor"), ("attention", "tensor"), ("layernorm", "hbm")):
                if fam not in fam_ms:
                    continue
                t = fam_ms[fam] * 1e-3
                if bound == "tensor":
                    ach, peak, unit = work[fam] * bs / t / 1e12, pk["bf16_sustained"], "TFLOP/s"
                else:
                    ach, peak, unit = work[fam] * bs / t / 1e9, pk["hbm"], "GB/s"
                families.append({"family": fam, "bound": bound, "achieved": ach, "peak": peak, "unit": unit,
                                 "frac": ach / peak, "ms_per_step": fam_ms[fam], "launches_per_step": fam_n[fam],
                                 "share_of_eager_step": fam_ms[fam] / eager_ms})
            families.append({"family": "other", "ms_per_step": fam_ms.get("other", 0.0),
                             "launches_per_step": fam_n.get("other", 0),
                             "share_of_eager_step": fam_ms.get("other", 0.0) / eager_ms,
                             "eager_step_ms": eager_ms, "graph_step_ms": ms / args.steps})
            gm = next(f for f in families if f["family"] == "gemm")
            traffic, tfile = measured_gemm_traffic()
            roofline = {"bound": "tensor", "kernel": "gemm_bf16_kernel / gemm2_bf16_kernel / patch_embed (tcgen05: all Linear "
                                                     "fwd/dgrad/wgrad + PatchEmbed)",
                        "achieved": gm["achieved"], "peak": pk["bf16_sustained"], "unit": "TFLOP/s", "frac": gm["frac"],
                        "frac_of_burst_peak": gm["achieved"] / pk["bf16_burst"],
                        "peak_source": pk["source"] + " (sustained cuBLAS bf16, kernel timed inside a long step)",
                        "traffic": traffic if (args.workload == "dino_vitb16" and bs == 128) else None,
                        "traffic_unit": f"bytes per launch (ncu dram read+write, profiles/{tfile})",
                        "algorithmic_bytes_per_launch": gemm_bytes_per_image(w) * bs / max(gm["launches_per_step"], 1),
                        "launches": gm["launches_per_step"] * n_f,
                        "avg_launch_us": gm["ms_per_step"] * 1e3 / max(gm["launches_per_step"], 1),
                        "share_of_step": gm["share_of_eager_step"],
                        "timed": f"CUDA events around each launch in an eager pass of {n_f} steps"}

    gpu_base = None
    if rank == 0 and world == 1 and not args.no_gpu_baseline:
        del trainer
        torch.cuda.empty_cache()
        gpu_base = gpu_eager_baseline(args, dev, bs)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args)

    if rank == 0:
        step_fl = step_flops_per_image(w) * bs
        kind = ("fine-tune (fwd+CE+bwd+SGD momentum 0.9)" if w["kind"] == "train" else
                "--lineareval (frozen backbone forward under no_grad + fc head [256,128,32,10] train step)")
        graph_on = not args.no_graph
        line = {
            "metric": "ViT-B/16 train images/sec", "value": value, "unit": "images/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {w['ctor']} {kind} 224x224 synthetic, random init",
                       "batch_per_gpu": bs, "global_batch": bs * world, "tokens": w["N"], "parallelism": f"dp{world}",
                       "input": {"u8": "uint8 NCHW pixels; ToTensor + Normalize on the GPU (vitk_normalize_u8), bf16 "
                                       "patch GEMM gathered by TMA from NCHW",
                                 "bf16": "bf16 normalised NCHW; patch GEMM gathered by TMA from NCHW",
                                 "f32": "fp32 normalised NCHW (the reference's loader output); tf32 patch GEMM gathered "
                                        "by TMA from NCHW"}[args.input],
                       "l2": "working set per step (>8 GB of activations) far exceeds the 126 MB L2; no explicit flush",
                       "numerics": "bf16 GEMM/attention operands, fp32 accumulate, fp32 residual stream + master weights",
                       "launch": ("eager launches" if not graph_on else
                                  "whole step captured in one CUDA graph" if world == 1 else
                                  "forward + loss + backward captured in one CUDA graph; gradient exchange and optimiser "
                                  "kernel launched after each replay"),
                       "grad_allreduce": {"none": "none (1 GPU)",
                                          "overlap": f"per-block NCCL all-reduces overlapped with backward, NCCL_MAX_CTAS={args.nccl_ctas}",
                                          "deferred": "one NCCL all-reduce over the fp32 gradient arena after backward",
                                          "bf16": "gradient arena cast to bf16 by one kernel, one NCCL all-reduce (bf16, "
                                                  "average), cast back into the fp32 arena",
                                          "split": "backward replayed as two CUDA graphs; the fp32 gradients of the last "
                                                   "two thirds of the blocks are all-reduced on a side stream by a "
                                                   f"communicator capped at {args.nccl_ctas} thread blocks while the second "
                                                   "graph runs; the remaining third is reduced after it",
                                          "registered": "one NCCL all-reduce over the fp32 gradient arena after backward; "
                                                        "the arena lives in ncclMemAlloc memory registered with the "
                                                        "communicator (zero-copy NVLS): "
                                                        + ("registered" if getattr(reducer, "arena_registered", False)
                                                           else "registration unavailable, plain buffer")}[dp_mode]},
            "step_tflops_per_gpu": step_fl / (ms / args.steps * 1e-3) / 1e12,
            "step_frac_of_bf16_peak": step_fl / (ms / args.steps * 1e-3) / 1e12 / pk["bf16_sustained"],
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline,
            "roofline_families": families, "gpu_eager_baseline": gpu_base, "cpu_baseline": cpu,
            "loss": float(loss),
        }
        _emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
