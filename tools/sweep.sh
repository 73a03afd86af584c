#!/bin/bash
# development aid: pipeline_only / value of bench.py for a few settings (environment knobs)
run() {
  python bench.py --steps 3 --warmup 2 --cpu-frames 0 --e2e-steps 1 --e2e-frames 240 --batch ${BATCH:-60} --streams ${STREAMS:-3} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['pipeline_only']['value']), {k: round(v*1e3,2) for k,v in list(d['kernel_ms_per_frame'].items())[:9]})"
}
APSE_CHAIN_GRID=1 run grid1
APSE_CHAIN_GRID=3 run grid3
APSE_CHAIN_GRID=4 run grid4
APSE_CHAIN_GRID=6 run grid6
