#!/bin/bash
B="python bench.py --steps 10 --cpu-frames 0 --e2e-steps 1"
for fpb in 20 10 7 5; do
  APSE_K1_FPB=$fpb $B 2>&1 | python tools/bsum.py fpb$fpb
done
APSE_K1_FPB=10 $B --streams 6 2>&1 | python tools/bsum.py fpb10_s6
APSE_K1_FPB=10 $B --streams 2 2>&1 | python tools/bsum.py fpb10_s2
