#!/usr/bin/env python3
"""Isolated timing of the fused attention kernel.  python tools/bench_attention.py [--tokens 197 --heads 12 --frames 256]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from dist_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--tokens", type=int, default=197)
ap.add_argument("--heads", type=int, default=12)
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--iters", type=int, default=10)
a = ap.parse_args()
qkv = (torch.randn(a.frames, a.tokens, 3 * a.heads * 64, device="cuda")).to(torch.bfloat16)
out = torch.empty(a.frames, a.tokens, a.heads * 64, device="cuda", dtype=torch.bfloat16)
call = ops.attention(qkv, out, a.frames, a.tokens, a.heads)
s = torch.cuda.current_stream()
for _ in range(3):
    call.launch(s.cuda_stream)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.iters):
    call.launch(s.cuda_stream)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.iters
print("attention frames=%d tokens=%d heads=%d: %.3f ms  %.1f TFLOP/s  %.1f GB/s" % (a.frames, a.tokens, a.heads, ms, call.flops / ms / 1e9, call.bytes / ms / 1e6))
