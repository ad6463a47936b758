#!/usr/bin/env python3
"""One eager (no CUDA graph) planned forward - the target of ``ncu`` captures.

    ncu --set full --clock-control none --import-source on --profile-from-start off -c 40 -o gpurun_out/layer0 python tools/run_once.py
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

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="b16_8x16")
ap.add_argument("--clips", type=int, default=32)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--names", default="", help="write the call names of the plan (launch order) to this JSON file")
a = ap.parse_args()
arch = DistArch(**WORKLOADS[a.workload]["arch"]).validate()
sd = synth.synth_state_dict(arch, seed=0)
eng = DistEngine(sd, arch, a.clips, precision=a.precision, text_features=synth.synth_text_features(arch.num_classes, arch.embed_dim))
eng.video.copy_(synth.synth_clips(2, arch).repeat(a.clips // 2, 1, 1, 1, 1))
torch.cuda.synchronize()
torch.cuda.profiler.start()          # ncu --profile-from-start off
eng.forward(use_graph=False)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", [c.name for c in eng.calls[:32]])
if a.names:
    import json
    json.dump([c.name for c in eng.calls], open(a.names, "w"))
