#!/usr/bin/env python3
"""Per-call timing of one planned forward (CUDA events around every launch), aggregated by call name.

    python tools/profile_plan.py [--workload b16_8x16] [--clips 32] [--precision bf16] [--out gpurun_out/plan.txt]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import WORKLOADS  # noqa: E402
from dist_b200.arch import DistArch  # noqa: E402
from dist_b200.engine import DistEngine  # noqa: E402
from dist_b200.utils import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="b16_8x16")
    ap.add_argument("--clips", type=int, default=32)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    arch = DistArch(**WORKLOADS[a.workload]["arch"]).validate()
    sd = synth.synth_state_dict(arch, seed=0)
    eng = DistEngine(sd, arch, a.clips, precision=a.precision, text_features=synth.synth_text_features(arch.num_classes, arch.embed_dim))
    eng.video.copy_(synth.synth_clips(2, arch).repeat(a.clips // 2, 1, 1, 1, 1))
    stream = torch.cuda.current_stream()
    agg = {}
    for rep in range(a.reps + 1):
        evs = []
        for c in eng.calls:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            c.launch(stream.cuda_stream)
            e1.record(stream)
            evs.append((c, e0, e1))
        torch.cuda.synchronize()
        if rep == 0:
            continue
        for c, e0, e1 in evs:
            s = agg.setdefault(c.name, [0.0, 0, 0, 0])
            s[0] += e0.elapsed_time(e1)
            s[1] += c.flops
            s[2] += c.bytes
            s[3] += 1
    tot = sum(s[0] for s in agg.values()) / a.reps
    lines = ["%-24s %5s %9s %7s %8s %8s" % ("call", "n", "ms/step", "share", "TFLOP/s", "GB/s")]
    for name, s in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        ms = s[0] / a.reps
        lines.append("%-24s %5d %9.3f %6.1f%% %8.1f %8.1f" % (name, s[3] // a.reps, ms, 100 * ms / tot, s[1] / s[0] / 1e9, s[2] / s[0] / 1e6))
    lines.append("total %.3f ms/step, %d launches, %.1f clips/s (sum of kernel times)" % (tot, eng.num_launches, a.clips / tot * 1e3))
    text = "\n".join(lines)
    print(text)
    if a.out:
        os.makedirs(os.path.dirname(a.out), exist_ok=True)
        open(a.out, "w").write(text + "\n")


if __name__ == "__main__":
    main()
