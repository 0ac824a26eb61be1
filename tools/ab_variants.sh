#!/bin/bash
# Kernel-variant A/B builds: tools/ab_variants.sh build   (here, cross-compiles)
#                            tools/ab_variants.sh run     (on the GPU box: K1b / K1 / K2 timings per variant)
# Variants are "name:EXTRA nvcc flags"; libraries land in qbold_vi_b200/variants/ (git-ignored, travels with gpurun).
set -e
cd "$(dirname "$0")/.."
VARIANTS=(
  "scalar5:-DQB_PACK_SMALL=0 -DQB_PACK_MID=0 -DQB_PACK_BIG=0"
  "packed5:"
  "packed4:-DQB_FWD_MIN_BLOCKS=4"
  "smallmid5:-DQB_PACK_BIG=0"
  "mid5:-DQB_PACK_SMALL=0 -DQB_PACK_BIG=0"
  "scalar4:-DQB_PACK_SMALL=0 -DQB_PACK_MID=0 -DQB_PACK_BIG=0 -DQB_FWD_MIN_BLOCKS=4"
)
if [ -n "$AB_VARIANTS" ]; then IFS=';' read -ra VARIANTS <<< "$AB_VARIANTS"; fi
mkdir -p qbold_vi_b200/variants
case "$1" in
build)
  for v in "${VARIANTS[@]}"; do
    name="${v%%:*}"; flags="${v#*:}"
    make -s -C qbold_vi_b200/csrc BUILD=../variants/$name/build OUT=../variants/libqbold_$name.so EXTRA="$flags" > /dev/null
    echo "$name: $(grep -A2 'k_forward_pairILb1' qbold_vi_b200/variants/$name/build/forward.ptxas.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ') | elbo: $(grep -A2 'k_elbo_pairILb1' qbold_vi_b200/variants/$name/build/elbo.ptxas.log | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
  done;;
run)
  for v in "${VARIANTS[@]}"; do
    name="${v%%:*}"
    echo "== $name"
    QBOLD_LIB=$PWD/qbold_vi_b200/variants/libqbold_$name.so python tools/kernel_bench.py --only "${AB_ONLY:-K}" --reps 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l); print('   %-45s %8.3f ms' % (d['kernel'], d['ms']))
    except Exception: print(l.rstrip())"
  done;;
esac
