#!/bin/bash
# On a GPU box: one `ncu --set full` capture of ONE frame per configuration -> gpurun_out/traffic.json (bench.py's `traffic` / `bound`
# fields, copied to profiles/traffic.json afterwards) and a text summary per configuration; the .ncu-rep files stay on the box (64 MiB limit).
#   usage: tools/capture_all.sh <tag> [configs...]
#   NCU_ARGS overrides `--set full` (c3: the volumes ncu saves and restores around every replay pass are ~15 GB, so its capture takes only the
#   metrics the summaries use: NCU_ARGS="--metrics <tools/traffic_from_ncu.py's list>")
tag=$1; shift
cfgs=${@:-c0 c1 c2 c3 c4}
mkdir -p gpurun_out
cp profiles/traffic.json gpurun_out/traffic.json 2>/dev/null
for c in $cfgs; do
  ncu ${NCU_ARGS:---set full} --clock-control none -k regex:^k_ -o /tmp/cap_$c -f python tools/profile_frame.py --config $c --iters 1 > /tmp/cap_$c.log 2>&1
  python tools/traffic_from_ncu.py /tmp/cap_$c.ncu-rep ${c}_k20 --out gpurun_out/traffic.json > gpurun_out/${tag}_${c}_launches.txt 2>&1
  python tools/ncu_summary.py /tmp/cap_$c.ncu-rep > gpurun_out/${tag}_${c}_ncu_full_summary.txt 2>&1
done
