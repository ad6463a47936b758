#!/usr/bin/env python3
"""Benchmark of the DiST video forward path (BASELINE.json: clips/sec, ViT-B/16 8+16f, 32 clips per GPU, bf16).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload b16_8x16|b16_32x64|l14_32x64]

One step = one forward of ``--clips`` (default 32) synthetic clips per GPU through the planned CUDA path
(CUDA-graph replay).  Rank 0 prints ONE JSON line:

  value      clips/s over all GPUs with the clips already resident in HBM, timed with CUDA events, max over ranks
  e2e        the same metric through the public API (``build_model(cfg)`` -> ``model({"video","texts"})``) with the
             clips in pinned HOST memory: H2D copy of the step's clips and D2H read of its class probabilities are
             inside the timed region
  roofline   the dominant kernel (tcgen05 GEMM): algorithmic FLOPs of its launches / their summed CUDA-event time
             (measured in a separate instrumented pass of the same plan), against MEASURED_PEAKS.json
  cpu_baseline  the oracle port (oracle/dist_oracle.py, fp32, all host threads) on a bounded sample of the workload
  kernels    time share per kernel family from the instrumented pass

``--impl reference`` times the CPU path only (rank 0; other ranks exit): the reference is PyTorch code that cannot
travel to the GPU box, so the arm runs the oracle port - the same torch-CPU tensor algebra - on the host cores.
Multi-GPU: one process per GPU (torchrun), clips sharded by rank, weights replicated, the only collective is the
all-gather of the per-clip class probabilities each step (runs/test.py:133 in the reference).
"""

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

from dist_b200.arch import DistArch  # noqa: E402
from dist_b200.utils import synth  # noqa: E402

WORKLOADS = {
    "b16_8x16": dict(arch=dict(), cfg="configs/projects/dist/ssv2/vit-b16-8+16f.yaml", label="DiST ViT-B/16 8+16f SSV2"),
    "b16_32x64": dict(arch=dict(frames=64, ada_layers=4, num_classes=400), cfg="configs/projects/dist/k400/vit-b16-32+64f.yaml",
                      label="DiST ViT-B/16 32+64f K400"),
    "b16_16x32": dict(arch=dict(frames=32), cfg="configs/projects/dist/ssv2/vit-b16-16+32f.yaml", label="DiST ViT-B/16 16+32f SSV2"),
    "l14_32x64": dict(arch=dict(width=1024, layers=24, patch=14, embed_dim=768, frames=64, s_patch=14, ada_layers=4,
                                num_classes=400, selected_layers=list(range(24))),
                      cfg="configs/projects/dist/k400/vit-l14-32+64f.yaml", label="DiST ViT-L/14 32+64f K400"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d["bf16_tflops_sustained"], source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(r[0]) for r in rows]
        busy = [s for s in sm if s > 0.5 * max(sm)] or sm
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": float(rows[0][1]), "power_w_max": max(float(r[2]) for r in rows),
                "reasons": reasons, "samples": len(rows)}


def cpu_reference(arch, sd, sample_clips, repeats):
    """Oracle port, fp32, all host threads; returns (clips/s, cores, description)."""
    from oracle import dist_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    clips = synth.synth_clips(sample_clips, arch, seed=1234, kind="structured")
    with torch.no_grad():
        dist_oracle.forward_arch(sd, clips[:1], arch, dtype=torch.float32)          # warm-up
        times = []
        for _ in range(repeats):
            t0 = time.perf_counter()
            dist_oracle.forward_arch(sd, clips, arch, dtype=torch.float32)
            times.append(time.perf_counter() - t0)
    best = statistics.median(times)
    return sample_clips / best, cores, "%d x forward of %d clip(s), fp32, median" % (repeats, sample_clips)


def run_reference_arm(args, arch, label):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sd = synth.synth_state_dict(arch, seed=0, init="reference")
    sample = 2 if arch.width < 1024 and arch.frames <= 16 else 1
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import dist_oracle
    clips = synth.synth_clips(sample, arch, seed=1234, kind="structured")
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 1))
    with torch.no_grad():
        for _ in range(warm):
            dist_oracle.forward_arch(sd, clips, arch, dtype=torch.float32)
        t0 = time.perf_counter()
        for _ in range(steps):
            dist_oracle.forward_arch(sd, clips, arch, dtype=torch.float32)
        el = time.perf_counter() - t0
    value = sample * steps / el
    line = {
        "impl": "reference", "metric": "clips/sec", "value": value, "unit": "clips/s", "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": el / steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": label + " forward, CPU oracle port of the reference path", "clips_per_step": sample},
        "cpu_baseline": {"value": value, "unit": "clips/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %d clip(s) per step, fp32, %d threads" % (steps, sample, cores)},
        "e2e": {"value": value, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def instrumented_pass(eng, reps=3):
    """Replay the plan eagerly with a CUDA event pair around every call; returns per-family stats."""
    stream = torch.cuda.current_stream()
    fam, by_name = {}, {}
    for rep in range(reps + 1):
        evs = []
        for c in eng.calls:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            c.launch(stream.cuda_stream)
            e1.record(stream)
            evs.append((c, e0, e1))
        torch.cuda.synchronize()
        if rep == 0:
            continue        # warm-up
        for c, e0, e1 in evs:
            key = family(c)
            f = fam.setdefault(key, dict(ms=0.0, flops=0, bytes=0, launches=0))
            f["ms"] += e0.elapsed_time(e1)
            f["flops"] += c.flops
            f["bytes"] += c.bytes
            f["launches"] += 1
            g = by_name.setdefault(c.name, dict(ms=0.0, flops=0, bytes=0, launches=0))
            g["ms"] += e0.elapsed_time(e1)
            g["flops"] += c.flops
            g["bytes"] += c.bytes
            g["launches"] += 1
    for f in list(fam.values()) + list(by_name.values()):
        f["ms"] /= reps
        f["flops"] //= reps
        f["bytes"] //= reps
        f["launches"] //= reps
    return fam, by_name


def family(call):
    n = call.name
    if n.startswith("patchify"):
        return "patchify"
    if n == "dist.tn":
        return "temporalnet"
    if "attention" in n or n.endswith(".attn"):
        return "attention" if n.startswith("vit") else "cross_attention"
    if n.endswith((".ln", ".ln_1", ".ln_2", ".ln_pre", ".ln_kv", ".ln_q", ".ln_out", "ln_post", ".stats")) or ".ln" in n:
        return "layernorm"
    if "cls" in n and ("rows" in n or n.endswith(".cls")) or n.startswith("ada.init"):
        return "rows_bcast"
    if n.endswith("cls_mean"):
        return "mean_rows"
    if n.startswith("head."):
        return "class_head"
    return "gemm_vit" if n.startswith("vit.") else "gemm_dist"


def run_train(args, arch, wl, world, rank, local, dev, light=False):
    """Fine-tuning step (frozen CLIP forward, DiST forward + backward, flat NCCL gradient all-reduce, AdamW).
    Returns the JSON line (rank 0) or None; ``light`` = the compact sub-record of the default run (no end-to-end leg)."""
    import torch.distributed as dist
    from dist_b200.train import TrainEngine
    sd = synth.synth_state_dict(arch, seed=0, init="reference")
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim)
    b = args.clips
    base = synth.synth_clips(min(b, 4), arch, seed=1234 + rank, kind="structured")
    clips = base.repeat((b + base.shape[0] - 1) // base.shape[0], 1, 1, 1, 1)[:b].contiguous()
    target = synth.synth_soft_targets(b, arch.num_classes, seed=99 + rank)
    eng = TrainEngine(sd, arch, b, device=dev, precision=args.precision, text_features=text, weight_decay=1e-4)
    eng.capture()
    d_clips, d_target = clips.to(dev), target.to(dev)
    lr = 3.2e-4                                        # BASE_LR 3.2e-5 x NEW_NET_LRMULT 10 (ssv2/vit-b16-16+32f.yaml:53-54)

    def timed(fn, steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        if world > 1:
            dist.barrier()
        ms = torch.tensor([max(e0.elapsed_time(e1), 0.0), wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    for i in range(args.warmup):
        eng.train_step(d_clips, d_target, lr)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    total_ms, _ = timed(lambda i: eng.train_step(d_clips, d_target, lr), args.steps)
    clocks = sampler.stop() if rank == 0 else None
    value = world * b * args.steps / (total_ms / 1e3)

    ems = None
    if not args.no_e2e:
        # end to end: pinned host clips + soft targets in, loss out, every step.  The clips of step i+1 cross PCIe on a copy stream
        # while step i computes (a pin_memory loader with .cuda(non_blocking=True), runs/train.py:87-89); every step's H2D copies
        # and the D2H read of the loss are inside the timed region.
        host = [clips.clone().pin_memory(), clips.flip(0).clone().pin_memory()]
        host_t = target.clone().pin_memory()
        loss_host = torch.empty(1).pin_memory()
        copy_stream = torch.cuda.Stream(dev)
        dev_buf = [torch.empty_like(d_clips) for _ in range(2)]
        dev_tgt = [torch.empty_like(d_target) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[i & 1])
                dev_buf[i & 1].copy_(host[i & 1], non_blocking=True)
                dev_tgt[i & 1].copy_(host_t, non_blocking=True)
                ready[i & 1].record(copy_stream)

        for ev in consumed:
            ev.record(torch.cuda.current_stream())
        prefetch(0)

        def e2e_step(i):
            prefetch(i + 1)
            torch.cuda.current_stream().wait_event(ready[i & 1])
            loss = eng.train_step(dev_buf[i & 1], dev_tgt[i & 1], lr)
            consumed[i & 1].record(torch.cuda.current_stream())
            loss_host.copy_(loss, non_blocking=False)

        for i in range(2):
            e2e_step(i)
        ems, ewall = timed(e2e_step, args.steps)
        ems = max(ems, ewall)
    if rank == 0:
        pk = peaks()
        fwd_fl, bwd_fl = sum(c.flops for c in eng.calls), sum(c.flops for c in eng.bwd)
        line = {
            "metric": "clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "%s fine-tuning step (frozen CLIP forward, DiST forward + backward, AdamW), %d synthetic clips per GPU, "
                                   "flat NCCL all-reduce of %.2f M trainable gradients" % (wl["label"], b, eng.pt.n_used / 1e6),
                       "clips_per_gpu": b, "precision": args.precision, "cuda_graph": "forward + backward", "optimizer": "AdamW lr 3.2e-4 wd 1e-4",
                       "gflop_per_clip": round((fwd_fl + bwd_fl) / b / 1e9, 1)},
            "model_tflops": value / world * (fwd_fl + bwd_fl) / b / 1e12,
            "e2e": None if ems is None else {"value": world * b * args.steps / (ems / 1e3), "unit": "clips/s",
                                             "h2d_bytes_per_step": int(clips.numel() * 4 + target.numel() * 4), "d2h_bytes_per_step": 4,
                                             "ms_per_step": ems / args.steps},
            "gpu_launches": (len(eng.calls) + len(eng.bwd) + len(eng.pack_calls) + 2) * args.steps,
            "launches_per_step": len(eng.calls) + len(eng.bwd) + len(eng.pack_calls) + 2, "loss": float(eng.loss), "clocks": clocks,
            "peaks": {"bf16_tflops_sustained": pk["tf_sustained"], "hbm_gbs": pk["hbm"]},
        }
    else:
        line = None
    del eng
    torch.cuda.empty_cache()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="dist_b200", choices=["dist_b200", "reference"])
    ap.add_argument("--workload", default="b16_8x16", choices=list(WORKLOADS))
    ap.add_argument("--clips", type=int, default=32, help="clips per GPU per step")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="headline workload only (skip the compact sub-records of the other BASELINE configs)")
    ap.add_argument("--mode", default="infer", choices=["infer", "train"],
                    help="train = one fine-tuning step per step (BASELINE configs[4]; use --workload b16_16x32)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    wl = WORKLOADS[args.workload]
    arch = DistArch(**wl["arch"]).validate()
    if args.impl == "reference":
        run_reference_arm(args, arch, wl["label"])
        return

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    import torch.distributed as dist
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.mode == "train":
        line = run_train(args, arch, wl, world, rank, local, dev)
    else:
        line = run_infer(args, arch, wl, world, rank, local, dev)
        if not args.no_extra and args.workload == "b16_8x16":
            # The other BASELINE.json configs as compact sub-records of the same run (same ranks, same clocks): the long-stack
            # B/16, ViT-L/14 (configs[2,3], the second half of the metric) and the fine-tuning step (configs[4]).
            import copy
            extra = {}
            for name, mode, steps, warm in (("b16_32x64", "infer", 10, 3), ("l14_32x64", "infer", 6, 3), ("b16_16x32", "train", 10, 3)):
                a2 = copy.copy(args)
                a2.workload, a2.steps, a2.warmup, a2.no_cpu_baseline = name, steps, warm, True
                wl2 = WORKLOADS[name]
                arch2 = DistArch(**wl2["arch"]).validate()
                fn = run_train if mode == "train" else run_infer
                sub = fn(a2, arch2, wl2, world, rank, local, dev, light=True)
                if sub is not None:
                    keep = ("value", "unit", "ms_per_step", "steps", "warmup", "dtype", "model_tflops", "e2e", "roofline", "roofline_hbm", "gemm_all",
                            "launches_per_step", "clocks", "loss")
                    extra[("train_" if mode == "train" else "") + name] = dict({k: sub[k] for k in keep if k in sub}, workload=sub["config"]["workload"])
            if line is not None:
                line["workloads"] = extra
    if rank == 0 and line is not None:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_infer(args, arch, wl, world, rank, local, dev, light=False):
    """Inference throughput of one workload.  Returns the JSON line (rank 0) or None; ``light`` = compact sub-record (no end-to-end
    leg, no CPU baseline)."""
    import torch.distributed as dist
    from dist_b200.engine import DistEngine
    sd = synth.synth_state_dict(arch, seed=0, init="reference")
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim)
    b = args.clips
    # clips: a few distinct structured clips tiled to the batch (generation cost only), different per rank
    base = synth.synth_clips(min(b, 4), arch, seed=1234 + rank, kind="structured")
    clips = base.repeat((b + base.shape[0] - 1) // base.shape[0], 1, 1, 1, 1)[:b].contiguous()

    eng = DistEngine(sd, arch, b, device=dev, precision=args.precision, text_features=text)
    eng.video.copy_(clips)
    eng.capture()
    gathered = torch.empty(world * b, eng.probs.shape[1], device=dev) if world > 1 else None
    # The logits gather of runs/test.py:133 runs on a side stream from a double-buffered copy of the step's probabilities, so that
    # its launch latency overlaps the next replay instead of sitting between two replays (round 1: 14.27 -> 14.46 ms at 8 GPUs).
    side = torch.cuda.Stream(dev) if world > 1 else None
    stage = [torch.empty_like(eng.probs) for _ in range(2)] if world > 1 else None
    staged_ev = [torch.cuda.Event() for _ in range(2)] if world > 1 else None
    gather_ev = [torch.cuda.Event() for _ in range(2)] if world > 1 else None
    counter = [0]

    def step():
        eng.graph.replay()
        if world > 1:
            i = counter[0] & 1
            main = torch.cuda.current_stream()
            if counter[0] >= 2:
                main.wait_event(gather_ev[i])            # the gather that last read this staging buffer
            stage[i].copy_(eng.probs, non_blocking=True)
            staged_ev[i].record(main)
            with torch.cuda.stream(side):
                side.wait_event(staged_ev[i])
                dist.all_gather_into_tensor(gathered, stage[i])
                gather_ev[i].record(side)
            counter[0] += 1

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    if world > 1:
        torch.cuda.current_stream().wait_stream(side)      # the last gathers are part of the timed region
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(ms.item())
    value = world * b * args.steps / (total_ms / 1e3)

    # ---- end to end through the public API, host buffers ----
    e2e = None
    e2e_u8 = None
    if not args.no_e2e:
        import dist_b200.models.base  # noqa: F401
        from dist_b200.config import Config
        from dist_b200.models.base.builder import build_model
        cfg = Config.from_file(os.path.join(ROOT, wl["cfg"]), ["NUM_GPUS", "1", "B200.PRECISION", args.precision])
        model, _ = build_model(cfg, gpu_id=local)
        model.eval()
        enc = model.module.backbone.base_encoder if hasattr(model, "module") else model.backbone.base_encoder
        enc.load_state_dict(sd, strict=True)
        text_dev = text.to(dev)
        out_host = [torch.empty(b, text.shape[0]).pin_memory() for _ in range(2)]
        out_done = [torch.cuda.Event() for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)

        def measure_e2e(host):
            """Pinned host clips -> device on a copy stream, one step ahead of the compute (what a pin_memory data loader with
            .cuda(non_blocking=True) does in runs/test.py:44-70); every step's H2D copy and D2H read are inside the timed region."""
            dev_buf = [torch.empty_like(host[0], device=dev) for _ in range(2)]
            ready = [torch.cuda.Event() for _ in range(2)]
            consumed = [torch.cuda.Event() for _ in range(2)]

            def prefetch(i):
                with torch.cuda.stream(copy_stream):
                    copy_stream.wait_event(consumed[i & 1])
                    dev_buf[i & 1].copy_(host[i & 1], non_blocking=True)
                    ready[i & 1].record(copy_stream)

            for ev in consumed:
                ev.record(torch.cuda.current_stream())
            prefetch(0)

            def e2e_step(i):
                prefetch(i + 1)
                torch.cuda.current_stream().wait_event(ready[i & 1])
                preds, _ = model({"video": dev_buf[i & 1], "texts": text_dev})
                consumed[i & 1].record(torch.cuda.current_stream())
                # D2H read of every step's result, stream-ordered behind the step; the host picks it up one step later (while the
                # GPU already runs the next step), the way a test loop accumulating per-video scores consumes it (runs/test.py:112-140)
                out_host[i & 1].copy_(preds, non_blocking=True)
                out_done[i & 1].record(torch.cuda.current_stream())
                if world > 1:
                    dist.all_gather_into_tensor(gathered, preds.contiguous())
                if i > 0:
                    out_done[(i - 1) & 1].synchronize()

            for i in range(args.warmup):
                e2e_step(i)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            e0.record()
            for i in range(args.steps):
                e2e_step(i)
            e1.record()
            torch.cuda.synchronize()
            wall = time.perf_counter() - t0
            if world > 1:
                dist.barrier()
            ems = torch.tensor([max(e0.elapsed_time(e1), wall * 1e3)], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(ems, op=dist.ReduceOp.MAX)
            return {"value": world * b * args.steps / (float(ems.item()) / 1e3), "unit": "clips/s",
                    "h2d_bytes_per_step": int(host[0].numel() * host[0].element_size()), "d2h_bytes_per_step": int(out_host[0].numel() * 4),
                    "ms_per_step": float(ems.item()) / args.steps}

        e2e = measure_e2e([clips.clone().pin_memory(), clips.flip(0).clone().pin_memory()])
        # the same call with decoded uint8 frames [b, T, H, W, 3] (normalisation fused into the patch-row kernel): a quarter of the bytes
        if not light:
            u8 = torch.randint(0, 256, (b, arch.frames, arch.resolution, arch.resolution, 3), dtype=torch.uint8)
            e2e_u8 = measure_e2e([u8.clone().pin_memory(), u8.flip(0).clone().pin_memory()])
        eng2 = next(e for k, (e, _, _) in enc._engines.items() if k[3] == "float")
    else:
        eng2 = eng

    if rank == 0:
        pk = peaks()
        eng_i = eng2 if e2e is not None else eng
        fam, by_name = instrumented_pass(eng_i)
        by_name_call = {c.name: c for c in eng_i.calls}
        tot = sum(f["ms"] for f in fam.values())
        gemm_ms = fam.get("gemm_vit", {}).get("ms", 0.0) + fam.get("gemm_dist", {}).get("ms", 0.0)
        gemm_fl = fam.get("gemm_vit", {}).get("flops", 0) + fam.get("gemm_dist", {}).get("flops", 0)
        achieved = gemm_fl / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        kernels = {k: {"ms": round(v["ms"], 4), "share": round(v["ms"] / tot, 4), "launches": v["launches"],
                       "tflops": round(v["flops"] / (v["ms"] / 1e3) / 1e12, 1) if v["ms"] > 0 else 0.0,
                       "gbs": round(v["bytes"] / (v["ms"] / 1e3) / 1e9, 1) if v["ms"] > 0 else 0.0} for k, v in sorted(fam.items())}
        # dominant kernel = the GEMM call site with the largest share of the step (FC1 of the ViT MLP)
        gemm_names = [n for n in by_name if family(by_name_call[n]) in ("gemm_vit", "gemm_dist")]
        top = max(gemm_names, key=lambda n: by_name[n]["ms"])
        t = by_name[top]
        top_tf = t["flops"] / (t["ms"] / 1e3) / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("%s/%s/%d" % (args.workload, top, b), {}).get("dram_bytes_per_launch")
        roofline = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel @ %s (%d launches per step)" % (top, t["launches"]),
                    "achieved": top_tf, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": top_tf / pk["tf_sustained"],
                    "traffic": traffic, "flop_per_launch": t["flops"] // max(t["launches"], 1), "us_per_launch": 1e3 * t["ms"] / max(t["launches"], 1),
                    "peak_source": pk["source"] + " sustained cuBLAS bf16 (kernel timed inside a long step); burst peak %.1f" % pk["tf_burst"]}
        # HBM-bound side: the fused TemporalNet launches (algorithmic bytes: fp32 stream in, fp32 stream + bf16 copy out, bf16 i2t term in)
        roofline_hbm = None
        if "dist.tn" in by_name:
            tn = by_name["dist.tn"]
            gbs = tn["bytes"] / (tn["ms"] / 1e3) / 1e9
            tn_traffic = None
            if os.path.exists(tpath):
                tn_traffic = json.load(open(tpath)).get("%s/dist.tn/%d" % (args.workload, b), {}).get("dram_bytes_per_launch")
            roofline_hbm = {"bound": "hbm", "kernel": "temporalnet_kernel @ dist.tn (%d launches per step)" % tn["launches"], "achieved": gbs,
                            "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"], "bytes_per_launch": tn["bytes"] // max(tn["launches"], 1),
                            "us_per_launch": 1e3 * tn["ms"] / max(tn["launches"], 1), "traffic": tn_traffic,
                            "note": "SIMT-issue / LSU bound, not HBM bound (DESIGN.md 4.7): the fraction is reported against the HBM roofline the north star names"}
        cpu = None
        if world > 1:
            cpu = {"value": None, "unit": "clips/s", "cores": 0, "kind": "port", "sample": "skipped: measured at N=1 only (the other ranks would idle in a barrier)"}
        elif not args.no_cpu_baseline and not light:
            heavy = arch.width >= 1024 or arch.frames > 16
            v, cores, sample = cpu_reference(arch, sd, 1 if heavy else 2, 1 if heavy else 3)
            cpu = {"value": v, "unit": "clips/s", "cores": cores, "kind": "port", "sample": sample}
        fl = arch.flops_per_clip()
        line = {
            "metric": "clips/sec", "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic",
            "config": {"workload": "%s inference, %d synthetic 224x224 clips per GPU, clip-sharded" % (wl["label"], b),
                       "clips_per_gpu": b, "precision": args.precision, "cuda_graph": True,
                       "l2": "per-step working set (%.1f GB of activations) exceeds the 126 MB L2; no explicit flush" % (
                           sum(t.numel() * t.element_size() for t in vars(eng).values() if torch.is_tensor(t)) / 1e9),
                       "gflop_per_clip": round(fl["total"] / 1e9, 1), "weights": "random init (reference distributions), seed 0"},
            "model_tflops": value / world * fl["total"] / 1e12,
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "gemm_all": {"launches": fam.get("gemm_vit", {}).get("launches", 0) + fam.get("gemm_dist", {}).get("launches", 0),
                         "achieved": achieved, "unit": "TFLOP/s", "frac": achieved / pk["tf_sustained"]},
            "cpu_baseline": cpu, "e2e": e2e, "e2e_uint8_frames": e2e_u8 if e2e is not None else None, "gpu_launches": eng.num_launches * args.steps, "launches_per_step": eng.num_launches,
            "kernels": kernels, "clocks": clocks,
        }
    else:
        line = None
    del eng, eng2
    if e2e is not None:
        del model, enc
    torch.cuda.empty_cache()
    return line


if __name__ == "__main__":
    main()
