#!/usr/bin/env python3
"""Isolated timing of the HBM-bound kernels at the shapes of the plan (32 clips): GB/s of algorithmic bytes against the measured
copy bandwidth (MEASURED_PEAKS.json).  Every case rotates over enough buffer sets to exceed the 126 MB L2.

    python tools/bench_hbm.py                     # all cases
    DISTB200_LIB=path/to/other.so python tools/bench_hbm.py     # A/B against another build of the library
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dist_b200 import ops

PEAK = 6556.2
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
dev = "cuda"
bf, f32 = torch.bfloat16, torch.float32


def timed(name, make, iters=40):
    sets = [make() for _ in range(4)]
    nbytes = sets[0].bytes
    need = max(1, int(300e6 // max(nbytes, 1)) + 1)
    while len(sets) < min(need, 12):
        sets.append(make())
    s = torch.cuda.current_stream().cuda_stream
    for c in sets:
        c.launch(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        sets[i % len(sets)].launch(s)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    gbs = nbytes / us / 1e3
    print("%-34s %8.1f us  %7.1f GB/s  %5.1f %% of %.0f" % (name, us, gbs, 100 * gbs / PEAK, PEAK))


def ln_case(rows, cols, dual=False):
    def make():
        x = torch.randn(rows, cols, device=dev)
        g, b = torch.ones(cols, device=dev), torch.zeros(cols, device=dev)
        y = torch.empty(rows, cols, device=dev, dtype=bf)
        if dual:
            y2 = torch.empty(rows, cols, device=dev, dtype=bf)
            return ops.layernorm(x, g, b, y, g2=g, b2=b, y2=y2)
        return ops.layernorm(x, g, b, y)
    return make


def rs_case(rows, cols):
    def make():
        x = torch.randn(rows, cols, device=dev).to(bf)
        st = torch.empty(rows, 2, device=dev)
        return ops.row_stats(x, st)
    return make


def patch_case(clips, T, p, n_sel, step):
    def make():
        v = torch.randn(clips, 3, T, 224, 224, device=dev)
        g = 224 // p
        ld = (3 * p * p + 7) // 8 * 8
        out = torch.empty(clips * n_sel * g * g, ld, device=dev, dtype=bf)
        return ops.patchify(v, out, clips, T, 224, 224, p, 0, step, n_sel, ld)
    return make


timed("row_stats bf16 [50432, 768]", rs_case(50432, 768))
timed("row_stats bf16 [65792, 1024]", rs_case(65792, 1024))
timed("layernorm [50432, 768] -> bf16", ln_case(50432, 768))
timed("layernorm [50432, 384] -> 2 x bf16", ln_case(50432, 384, True))
timed("layernorm [100352, 96] -> bf16", ln_case(100352, 96))
timed("layernorm [65792, 1024] -> bf16", ln_case(65792, 1024))
timed("patchify p=16 32 clips x 16 frames", patch_case(32, 16, 16, 16, 1), iters=12)
timed("patchify p=14 8 clips x 64 frames", patch_case(8, 64, 14, 64, 1), iters=12)
