#!/bin/bash
# Run on the GPU box (gpurun -- 'bash tools/refresh_profiles_r2.sh'): regenerates the round-2 measurements that profiles/ summarises.
# Outputs land in gpurun_out/; `python tools/summarize_profiles.py --round r2` (build container) turns them into profiles/*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 30 --warmup 5 > $O/bench_b16_8x16.json 2> $O/bench_b16_8x16.err
timeout 300 python tools/profile_plan.py --out $O/plan_b16_8x16.txt > /dev/null 2>&1
timeout 300 python tools/profile_plan.py --workload l14_32x64 --clips 32 --out $O/plan_l14_32x64.txt > /dev/null 2>&1
timeout 300 python tools/profile_plan.py --workload b16_32x64 --clips 32 --out $O/plan_b16_32x64.txt > /dev/null 2>&1
timeout 300 python tools/profile_train.py --out $O/plan_train_b16_16x32.txt > /dev/null 2>&1
kill $SMI
# launch list of one forward (cold-cache, serialised: compare SHARES with the bench line's "kernels")
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extra > /dev/null 2>&1
# full-set capture: first ViT layer + first DiST layer + the head of the second ViT layer (launches 5..22 of the eager plan)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip 5 -c 17 -o /tmp/layer0 -f \
    python tools/run_once.py --names $O/layer0_names.json > /dev/null 2>&1
ncu -i /tmp/layer0.ncu-rep --page raw --csv > $O/layer0_raw.csv 2>/dev/null
# the same for the first ViT layer of ViT-L/14 32+64f (FC1 / attention at 257 tokens, 32 clips)
timeout 900 ncu --set full --clock-control none --profile-from-start off --launch-skip 5 -c 7 -o /tmp/l14 -f \
    python tools/run_once.py --workload l14_32x64 --clips 32 --names $O/l14_names.json > /dev/null 2>&1
ncu -i /tmp/l14.ncu-rep --page raw --csv > $O/l14_raw.csv 2>/dev/null
ls -la $O
