#!/usr/bin/env python3
"""Timeline of the fused attention kernel (a -DDISTB200_ATT_PROBES build): clock64 stamps per tile of the first CTA.

    python tools/trace_attention.py --build      # build container: dist_b200/libdistb200_attprobe.so
    python tools/trace_attention.py [--tokens 197 --heads 12 --frames 256]
"""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
LIB = os.path.join(ROOT, "dist_b200", "libdistb200_attprobe.so")
SLOTS = {0: "S.wait", 1: "S.issue", 2: "PV.issue", 4: "sm.top", 5: "sm.s_full", 6: "sm.max_done", 7: "sm.p_done", 8: "sm.o_full", 9: "sm.stored"}

ap = argparse.ArgumentParser()
ap.add_argument("--build", action="store_true")
ap.add_argument("--tokens", type=int, default=197)
ap.add_argument("--heads", type=int, default=12)
ap.add_argument("--frames", type=int, default=256)
ap.add_argument("--lib", default=LIB)
ap.add_argument("--first", type=int, default=0)
ap.add_argument("--count", type=int, default=12)
ap.add_argument("--cta", type=int, default=0)
a = ap.parse_args()
if a.build:
    from dist_b200 import build as b
    subprocess.run(["nvcc"] + b.NVCC_FLAGS + ["-DDISTB200_ATT_PROBES", "-o", LIB] + b.sources(), check=True)
    print(LIB)
    sys.exit(0)
os.environ["DISTB200_LIB"] = a.lib
os.environ["DISTB200_ALLOW_STALE"] = "1"
import torch
from dist_b200 import ops
L = ops.lib()
L.distb200_debug_att_trace.argtypes = [ctypes.c_void_p]
qkv = torch.randn(a.frames, a.tokens, 3 * a.heads * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(a.frames, a.tokens, a.heads * 64, device="cuda", dtype=torch.bfloat16)
call = ops.attention(qkv, out, a.frames, a.tokens, a.heads)
s = torch.cuda.current_stream()
for _ in range(3):
    call.launch(s.cuda_stream)
torch.cuda.synchronize()
trace = torch.zeros(148 * 64 * 16, dtype=torch.int64, device="cuda")
L.distb200_debug_att_trace(trace.data_ptr())
call.launch(s.cuda_stream)
torch.cuda.synchronize()
L.distb200_debug_att_trace(None)
tr = trace.cpu().view(148, 64, 16)
rows = tr[a.cta]
t0 = int(rows[rows > 0].min())
for g in range(a.first, min(64, a.first + a.count)):
    if int(rows[g].max()) == 0:
        break
    ev = sorted((int(rows[g][k]) - t0, n) for k, n in SLOTS.items() if int(rows[g][k]) > 0)
    print("tile %2d (slot %d): " % (g, g % 2) + "  ".join("%s %d" % (n, v) for v, n in ev))
