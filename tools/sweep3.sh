#!/bin/bash
B="python bench.py --steps 10 --cpu-frames 0 --e2e-steps 1"
for nreg in 64 56 48; do for g in 4 2 1; do
  APSE_K1_NREG=$nreg APSE_CHAIN_GRID=$g $B 2>&1 | python tools/bsum.py nreg${nreg}_grid$g | cut -c1-60
done; done
