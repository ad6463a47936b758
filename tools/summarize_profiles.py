#!/usr/bin/env python3
"""Turn the raw outputs of tools/refresh_profiles.sh (gpurun_out/) into the tracked summaries under profiles/.

    python tools/summarize_profiles.py [--round r1]
"""
import argparse
import csv
import json
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = os.path.join(ROOT, "gpurun_out")
P = os.path.join(ROOT, "profiles")

LAYER0 = ["vit.ln_1", "vit.qkv", "vit.attention", "vit.out_proj", "vit.ln_2.stats", "vit.fc1", "vit.fc2", "dist.tn.ln", "dist.tn.conv_t",
          "dist.tn.conv_s", "dist.input_linear", "dist.i2t", "dist.int.stats", "dist.int.fc", "dist.int.t_conv",
          "vit.ln_1.stats", "vit.qkv#1", "vit.attention#1"]          # the capture runs three launches into the second ViT block
METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed.avg.per_cycle_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_tensor.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__block_size",
           "launch__cluster_size", "sm__warps_active.avg.pct_of_peak_sustained_active"]


def launches(rnd):
    path = os.path.join(G, "launches.csv")
    if not os.path.exists(path):
        return
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    head = next(r for r in rows if "Kernel Name" in r)
    ik, iv, im = head.index("Kernel Name"), head.index("Metric Value"), head.index("Metric Name")
    data = [(r[ik], float(r[iv].replace(",", ""))) for r in rows if r is not head and len(r) > iv and r[im] == "gpu__time_duration.sum"]
    shutil.copy(path, os.path.join(P, rnd + "_launches.csv"))
    bench = json.load(open(os.path.join(G, "bench_b16_8x16.json")))
    n = bench["launches_per_step"]
    last = data[-n:]
    agg = {}
    for k, v in last:
        k = k.split("(")[0].replace("distb200::", "").replace("<unnamed>::", "").strip()
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    unit_ns = tot > 1e6
    with open(os.path.join(P, rnd + "_launches_summary.txt"), "w") as f:
        f.write("# ncu launch list of one forward (last %d launches of `ncu --metrics gpu__time_duration.sum --clock-control none -c 700 "
                "python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline`)\n" % n)
        f.write("# times are cold-cache / serialised under the profiler: compare SHARES with bench.py's \"kernels\" table, not absolutes\n")
        f.write("%-52s %5s %12s %7s\n" % ("kernel", "n", "sum us", "share"))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write("%-52s %5d %12.1f %6.1f%%\n" % (k[:52], a[0], a[1] / (1e3 if unit_ns else 1), 100 * a[1] / tot))


def layer0(rnd):
    path = os.path.join(G, "layer0_raw.csv")
    if not os.path.exists(path):
        return
    r = list(csv.reader(open(path)))
    h, units, rows = r[0], r[1], r[2:]
    traffic = {}
    with open(os.path.join(P, rnd + "_ncu_full_summary.txt"), "w") as f:
        f.write("ncu --set full --clock-control none --import-source on: first ViT layer + first DiST layer of one eager forward\n"
                "(B/16 8+16f, 32 clips; tools/run_once.py; one launch each, caches flushed by ncu between replays)\n\n")
        for name, row in zip(LAYER0, rows):
            f.write("== %s   %s\n" % (name, row[h.index("Kernel Name")][:90]))
            for m in METRICS:
                if m in h:
                    f.write("   %-72s %s %s\n" % (m, row[h.index(m)], units[h.index(m)]))
            rd, wr = float(row[h.index("dram__bytes_read.sum")]), float(row[h.index("dram__bytes_write.sum")])
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            tb = rd * scale[units[h.index("dram__bytes_read.sum")]] + wr * scale[units[h.index("dram__bytes_write.sum")]]
            traffic["b16_8x16/%s/32" % name] = {"dram_bytes_per_launch": int(tb), "source": "profiles/%s_ncu_full_summary.txt (ncu --set full, one launch)" % rnd}
            f.write("\n")
    json.dump(traffic, open(os.path.join(P, "traffic.json"), "w"), indent=1)


def capture(rnd, raw_name, names_name, skip, workload, clips, out_name, title):
    """ncu --set full capture of consecutive launches of the eager plan: one block of metrics per call site (names from run_once.py)."""
    path, npath = os.path.join(G, raw_name), os.path.join(G, names_name)
    if not (os.path.exists(path) and os.path.exists(npath) and os.path.getsize(path) > 0):
        return
    names = json.load(open(npath))[skip:]
    r = list(csv.reader(open(path)))
    h, units, rows = r[0], r[1], r[2:]
    tpath = os.path.join(P, "traffic.json")
    traffic = json.load(open(tpath)) if os.path.exists(tpath) else {}
    scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
    seen = {}
    with open(os.path.join(P, out_name), "w") as f:
        f.write(title + "\n\n")
        for name, row in zip(names, rows):
            seen[name] = seen.get(name, 0) + 1
            tag = name if seen[name] == 1 else "%s#%d" % (name, seen[name] - 1)
            f.write("== %s   %s\n" % (tag, row[h.index("Kernel Name")][:100]))
            for m in METRICS:
                if m in h:
                    f.write("   %-78s %s %s\n" % (m, row[h.index(m)], units[h.index(m)]))
            rd, wr = float(row[h.index("dram__bytes_read.sum")]), float(row[h.index("dram__bytes_write.sum")])
            tb = rd * scale[units[h.index("dram__bytes_read.sum")]] + wr * scale[units[h.index("dram__bytes_write.sum")]]
            if seen[name] == 1:
                traffic["%s/%s/%d" % (workload, name, clips)] = {"dram_bytes_per_launch": int(tb), "source": "profiles/%s (ncu --set full, one launch)" % out_name}
            f.write("\n")
    json.dump(traffic, open(tpath, "w"), indent=1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--round", default="r1")
    a = ap.parse_args()
    os.makedirs(P, exist_ok=True)
    for name in ("bench_b16_8x16.json", "bench_l14_32x64.json", "bench_b16_32x64.json", "bench_train_b16_16x32.json", "plan_b16_8x16.txt",
                 "plan_l14_32x64.txt", "plan_b16_32x64.txt", "plan_train_b16_16x32.txt"):
        src = os.path.join(G, name)
        if os.path.exists(src) and os.path.getsize(src) > 0:
            shutil.copy(src, os.path.join(P, a.round + "_" + name))
    launches(a.round)
    if a.round == "r1":
        layer0(a.round)
    else:
        capture(a.round, "layer0_raw.csv", "layer0_names.json", 5, "b16_8x16", 32, a.round + "_ncu_full_summary.txt",
                "ncu --set full --clock-control none --import-source on: first ViT layer, first DiST layer and the head of the second ViT layer of one "
                "eager forward\n(B/16 8+16f, 32 clips; tools/run_once.py; one launch each, caches flushed by ncu between replays)")
        capture(a.round, "l14_raw.csv", "l14_names.json", 5, "l14_32x64", 32, a.round + "_ncu_full_l14_summary.txt",
                "ncu --set full --clock-control none: first ViT layer of ViT-L/14 32+64f, 32 clips (257 tokens; tools/run_once.py --workload l14_32x64)")
        for extra in ("clocks.csv",):
            src = os.path.join(G, extra)
            if os.path.exists(src):
                shutil.copy(src, os.path.join(P, a.round + "_" + extra))
    print("profiles/:", sorted(os.listdir(P)))


if __name__ == "__main__":
    main()
