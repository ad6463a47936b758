#!/usr/bin/env python3
"""Per-call timing of one planned fine-tuning step (forward + backward + update), aggregated by call name.

    python tools/profile_train.py [--frames 32] [--clips 32] [--precision bf16] [--out gpurun_out/train_plan.txt]
"""
import argparse
import os
import re
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dist_b200.arch import DistArch  # noqa: E402
from dist_b200.train import TrainEngine  # noqa: E402
from dist_b200.utils import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--clips", type=int, default=32)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    arch = DistArch(frames=a.frames).validate()
    sd = synth.synth_state_dict(arch, seed=0)
    text = synth.synth_text_features(arch.num_classes, arch.embed_dim)
    eng = TrainEngine(sd, arch, a.clips, precision=a.precision, text_features=text)
    clips = synth.synth_clips(2, arch).repeat(a.clips // 2, 1, 1, 1, 1).cuda()
    target = synth.synth_soft_targets(a.clips, arch.num_classes).cuda()
    print("params %.2f M, device memory %.1f GB" % (eng.pt.n_used / 1e6, torch.cuda.memory_allocated() / 2**30))
    stream = torch.cuda.current_stream()
    agg = {}
    for rep in range(a.reps + 1):
        eng.video.copy_(clips)
        eng.target.copy_(target)
        eng.pt.g.zero_()
        eng.loss.zero_()
        evs = []
        for phase, calls in (("fwd", eng.calls), ("bwd", eng.bwd)):
            for c in calls:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                c.launch(stream.cuda_stream)
                e1.record(stream)
                evs.append((phase, c, e0, e1))
        torch.cuda.synchronize()
        if rep == 0:
            continue
        for phase, c, e0, e1 in evs:
            key = phase + ":" + re.sub(r"(dist|ada)\d+\.", r"\1.", c.name.replace("bwd.", ""))
            s = agg.setdefault(key, [0.0, 0, 0, 0])
            s[0] += e0.elapsed_time(e1)
            s[1] += c.flops
            s[2] += c.bytes
            s[3] += 1
    tot = sum(s[0] for s in agg.values()) / a.reps
    lines = ["%-36s %5s %9s %7s %8s %8s" % ("call", "n", "ms/step", "share", "TFLOP/s", "GB/s")]
    for name, s in sorted(agg.items(), key=lambda kv: -kv[1][0])[:60]:
        ms = s[0] / a.reps
        lines.append("%-36s %5d %9.3f %6.1f%% %8.1f %8.1f" % (name, s[3] // a.reps, ms, 100 * ms / tot, s[1] / s[0] / 1e9, s[2] / s[0] / 1e6))
    fwd = sum(s[0] for k, s in agg.items() if k.startswith("fwd")) / a.reps
    lines.append("forward %.3f ms, backward %.3f ms, %d + %d launches (sum of kernel times)" % (fwd, tot - fwd, len(eng.calls), len(eng.bwd)))
    # whole step with the CUDA graph and the update
    eng.capture()
    for _ in range(2):
        eng.train_step(clips, target, 1e-4)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 5
    for _ in range(n):
        eng.train_step(clips, target, 1e-4)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / n * 1e3
    lines.append("train_step (graph + all-reduce-free update): %.2f ms/step, %.1f clips/s, loss %.4f" % (ms, a.clips / ms * 1e3, float(eng.loss)))
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
