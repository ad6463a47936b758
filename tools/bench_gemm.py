#!/usr/bin/env python3
"""Isolated timing of the ViT GEMM shapes (CUDA events, L2 flushed between iterations by the operand sizes).

    python tools/bench_gemm.py [--which fc1,qkv,proj,fc2] [--m 50432] [--iters 10]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dist_b200 import ops  # noqa: E402

SHAPES = {  # name: (N, K, act, residual, out dtype, out2)
    "qkv": (2304, 768, False, False, torch.bfloat16, False),
    "proj": (768, 768, False, True, torch.float32, False),
    "fc1": (3072, 768, True, False, torch.bfloat16, False),
    "fc2": (768, 3072, False, True, torch.float32, True),
    "proj2": (768, 768, False, True, torch.float32, True),          # the planned out_proj: fp32 stream in place + bf16 copy
    "proj2_nores": (768, 768, False, False, torch.float32, True),   # same stores, no residual read
    "proj_bf": (768, 768, False, False, torch.bfloat16, False),     # bf16 store only
    "plain": (3072, 768, False, False, torch.bfloat16, False),
    "nobias": (3072, 768, False, False, torch.bfloat16, False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--which", default="qkv,proj,fc1,fc2")
    ap.add_argument("--m", type=int, default=50432)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--block-n", type=int, default=0)
    a = ap.parse_args()
    dev = "cuda"
    for name in a.which.split(","):
        N, K, act, res, odt, out2 = SHAPES[name]
        A = (torch.randn(a.m, K, device=dev) * 0.5).to(torch.bfloat16)
        W = (torch.randn(N, K, device=dev) / K ** 0.5).to(torch.bfloat16)
        bias = None if name == "nobias" else torch.randn(N, device=dev)
        h = torch.randn(a.m, N, device=dev) if res else None
        out = h if res else torch.empty(a.m, N, device=dev, dtype=odt)
        o2 = torch.empty(a.m, N, device=dev, dtype=torch.bfloat16) if out2 else None
        call = ops.gemm(A, W, N, K, bias=bias, res=h, ld_res=N, out=out, ld_out=N, out2=o2, ld_out2=N,
                        act=ops.ACT_QUICKGELU if act else ops.ACT_NONE, block_n=a.block_n)
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
        print("%-7s M=%d N=%d K=%d: %.3f ms  %.1f TFLOP/s  %.1f GB/s" % (name, a.m, N, K, ms, call.flops / ms / 1e9, call.bytes / ms / 1e6))


if __name__ == "__main__":
    main()
