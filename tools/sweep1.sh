#!/bin/bash
# co-scheduling sweep: preprocess register cap x stream priorities
B="python bench.py --steps 10 --cpu-frames 0 --e2e-steps 1"
for nreg in 64 56 48; do
  APSE_K1_NREG=$nreg $B 2>&1 | python tools/bsum.py nreg$nreg
done
APSE_K1_NREG=48 APSE_CHAIN_PRIO=0 $B 2>&1 | python tools/bsum.py nreg48_eqprio
APSE_K1_NREG=48 APSE_CHAIN_PRIO=0 APSE_PRE_PRIO=-1 $B 2>&1 | python tools/bsum.py nreg48_prehigh
APSE_K1_NREG=64 APSE_CHAIN_PRIO=0 $B 2>&1 | python tools/bsum.py nreg64_eqprio
APSE_K1_NREG=48 $B --streams 4 2>&1 | python tools/bsum.py nreg48_s4
APSE_K1_NREG=48 $B --streams 2 --batch 120 2>&1 | python tools/bsum.py nreg48_s2b120
