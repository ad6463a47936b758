#!/usr/bin/env python3
"""Timeline trace and bottleneck probes of the fused TemporalNet kernel (a -DDISTB200_TN_TRACE build of the library).

    python tools/trace_temporalnet.py --build            # build container: writes dist_b200/libdistb200_trace.so
    python tools/trace_temporalnet.py [--grid 14 --frames 16 --clips 32]     # GPU box

Per probe setting: time of the launch (CUDA events, L2 flushed between launches) and, for the plain run, the clock64 stamps of
the first CTAs per iteration (cycles relative to the CTA's first stamp).
"""
import argparse
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TRACE_LIB = os.path.join(ROOT, "dist_b200", "libdistb200_trace.so")

SLOTS = {0: "ln.wait0", 1: "ln.wait1", 2: "ln.done", 28: "ln.fenced", 3: "ln.arrived", 4: "mma.w311", 5: "mma.i311", 6: "mma.w133", 7: "mma.i133",
         8: "epi.top", 9: "epi.c311", 10: "epi.c133", 14: "e1.math", 15: "e1.fenced", 11: "e1.arrived",
         16: "c0.tmem", 17: "c0.staged", 18: "c0.pref", 19: "c0.stored", 20: "c1.tmem", 21: "c1.staged", 22: "c1.pref", 23: "c1.stored",
         24: "c2.tmem", 12: "acc_rel", 25: "c2.staged", 27: "c2.stored", 13: "out_done"}


def build():
    from dist_b200 import build as b
    cmd = ["nvcc"] + b.NVCC_FLAGS + ["-DDISTB200_TN_TRACE", "-o", TRACE_LIB] + b.sources()
    subprocess.run(cmd, check=True)
    print(TRACE_LIB)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--build", action="store_true")
    ap.add_argument("--grid", type=int, default=14)
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--clips", type=int, default=32)
    ap.add_argument("--alpha", type=int, default=2)
    ap.add_argument("--no-u", action="store_true")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--plain", action="store_true", help="the production library, plain launches only (for ncu)")
    a = ap.parse_args()
    if a.build:
        return build()
    if not a.plain:
        os.environ["DISTB200_LIB"] = TRACE_LIB
    import torch
    from dist_b200 import ops
    L = ops.lib()
    if not a.plain:
        L.distb200_debug_tn_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    C, g, T, B, al = 96, a.grid, a.frames, a.clips, a.alpha
    P = g * g
    dev = "cuda"
    x = torch.randn(B, T, P, C, device=dev)
    u = None if a.no_u else torch.randn(B, T // al, P, C, device=dev).to(torch.bfloat16)
    gam, bet, b1, b2 = torch.ones(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    w1 = (torch.randn(3, C, C, device=dev) / 17).to(torch.bfloat16)
    w2 = (torch.randn(9, C, C, device=dev) / 29).to(torch.bfloat16)
    out = torch.empty(B, T, P, C, device=dev)
    out2 = torch.empty(B * T // al * (P + 1), 1984, device=dev, dtype=torch.bfloat16)
    call = ops.temporalnet(x, gam, bet, w1, b1, w2, b2, clips=B, frames=T, grid=g, u=u, alpha=al, out=out, out2=out2[:, 768:], ld_out2=1984,
                           out2_gdiv=al, out2_cstep=C, out2_gstride=P + 1, out2_roff=1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    s = torch.cuda.current_stream()
    trace = torch.zeros(148 * 64 * 32, dtype=torch.int64, device=dev)
    algo = B * T * P * C * (4 + 4 + 2) + (0 if u is None else B * T // al * P * C * 2)

    def timed(dbg, with_trace=False):
        if not a.plain:
            L.distb200_debug_tn_trace(trace.data_ptr() if with_trace else None, dbg)
        ts = []
        for i in range(a.reps + 2):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s)
            call.launch(s.cuda_stream)
            e1.record(s)
            torch.cuda.synchronize()
            if i >= 2:
                ts.append(e0.elapsed_time(e1) * 1e3)
        ts.sort()
        return ts[len(ts) // 2], ts[0]

    print("TemporalNet C=%d g=%d T=%d clips=%d alpha=%d u=%s: algorithmic %.1f MB" % (C, g, T, B, al, u is not None, algo / 1e6))
    if a.plain:
        med, best = timed(0)
        print("  production library: %8.1f us (best %.1f)  %6.0f GB/s algorithmic" % (med, best, algo / med / 1e3))
        return
    for dbg, label in ((0, "plain"), (1, "no MMA issue"), (2, "no LN loads"), (4, "no residual loads / stores"), (8, "no z epilogue"),
                       (16, "CTA-scope arrives"), (32, "no L2 prefetch"), (1 | 2 | 4 | 8, "barriers only"), (2 | 4, "no global traffic"), (1 | 8, "memory only")):
        med, best = timed(dbg)
        print("  dbg %2d %-28s %8.1f us (best %.1f)  %6.0f GB/s algorithmic" % (dbg, label, med, best, algo / med / 1e3))
    timed(0, with_trace=True)
    tr = trace.cpu().view(148, 64, 32)
    for cta in (0, 1):
        rows = tr[cta]
        nz = rows[rows > 0]
        if nz.numel() == 0:
            continue
        t0 = int(nz.min())
        print("CTA %d (cycles since its first stamp)" % cta)
        for it in range(8):
            if int(rows[it].max()) == 0:
                break
            ev = sorted((int(rows[it][k]) - t0, n) for k, n in SLOTS.items() if int(rows[it][k]) > 0)
            print("  it %d: " % it + "  ".join("%s %d" % (n, v) for v, n in ev))


if __name__ == "__main__":
    main()
