#!/bin/bash
# Run on the GPU box (gpurun -- 'bash tools/refresh_profiles.sh'): regenerates every measurement that profiles/ summarises.
# Outputs land in gpurun_out/; `python tools/summarize_profiles.py` (build container) turns them into profiles/*.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
python bench.py --steps 30 --warmup 5 > $O/bench_b16_8x16.json 2> $O/bench_b16_8x16.err
python tools/profile_plan.py --out $O/plan_b16_8x16.txt > /dev/null 2>&1
python tools/profile_plan.py --workload l14_32x64 --clips 32 --out $O/plan_l14_32x64.txt > /dev/null 2>&1
python tools/profile_plan.py --workload b16_32x64 --clips 32 --out $O/plan_b16_32x64.txt > /dev/null 2>&1
python bench.py --workload l14_32x64 --clips 32 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_l14_32x64.json 2> $O/bench_l14.err
python bench.py --workload b16_32x64 --clips 32 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > $O/bench_b16_32x64.json 2> $O/bench_b16_32x64.err
python bench.py --mode train --workload b16_16x32 --steps 10 --warmup 3 > $O/bench_train_b16_16x32.json 2> $O/bench_train.err
python tools/profile_train.py --out $O/plan_train_b16_16x32.txt > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
# full-set capture of the first ViT layer and the first DiST layer (launches 5..22 of the plan)
ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip 5 -c 18 -o /tmp/layer0 -f \
    python tools/run_once.py > /dev/null 2>&1
ncu -i /tmp/layer0.ncu-rep --page raw --csv > $O/layer0_raw.csv 2>/dev/null
ls -la $O
