#!/bin/bash
# sweep the k-stride kernel's CTA size (occupancy) on config 2; prints block, distribution, GB/s
for b in 1024 640 512 768 448; do
  FSM_B200_KSTRIDE_BLOCK=$b VARIANTS=kstride PADS=4 python tools/bench_variants.py 2>&1 | tail -2 | sed "s/^/block=$b /"
done
