#!/bin/bash
# A/B of the cost-volume kernels' tuning knobs on one configuration: prints the K1a / K1b event times per variant.
#   usage: tools/k1_ab.sh <config> "<ENV=.. ENV=..>" "<...>" ...
cfg=$1; shift
for v in "$@"; do
  env $v python tools/profile_frame.py --config $cfg --iters 3 --avg 10 2>&1 | tr ' ' '\n' | grep -E "k_ad|k_box|total" | tr '\n' ' '
  echo " <- [$v]"
done
