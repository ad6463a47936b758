#!/bin/bash
# Bottleneck probes of the tcgen05 GEMM (run on the GPU box): rebuilds the library with -DDISTB200_GEMM_PROBES and
# prints the per-call-site times of one planned forward for each probe mask (see gemm_tcgen05.cu: 1 = no epilogue
# memory traffic, 2 = no MMA issue, 4 = no A loads, 8 = no B loads, 16 = no TMEM reads, 32 = one k-iteration per tile).
#   bash tools/probe_gemm.sh "0 1 2 4 8 15 63"
set -e
cd "$(dirname "$0")/.."
DISTB200_NVCC_EXTRA=-DDISTB200_GEMM_PROBES python -m dist_b200.build --force > /dev/null
for dbg in ${1:-0 1 2 4 8 12 15 63}; do
  echo "== DISTB200_GEMM_DBG=$dbg"
  DISTB200_GEMM_DBG=$dbg python tools/profile_plan.py 2>&1 | grep -E "dist\.(tn.conv|input|t2i |i2t|int\.)|vit\.(out_proj|fc|qkv)" | awk '{printf "%s=%s ", $1, $3} END {print ""}'
done
python -m dist_b200.build --force > /dev/null     # restore the production build
