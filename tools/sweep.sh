#!/bin/bash
# development aid: pipeline_only / value of bench.py for a few settings (environment knobs)
run() {
  python bench.py --steps 3 --warmup 2 --cpu-frames 0 --e2e-steps 1 --e2e-frames 240 --batch ${BATCH:-60} --streams ${STREAMS:-3} 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$1', round(d['value']), round(d['pipeline_only']['value']), round(1e3*d['kernel_ms_per_frame']['k_preprocess_fused'],2))"
}
run base_48_4
APSE_K1B_NREG=40 run r40_s4
APSE_K1B_NREG=40 APSE_K1B_STAGES=3 run r40_s3
APSE_K1B_NREG=40 APSE_K1B_STAGES=2 run r40_s2
APSE_K1B_STAGES=3 run r48_s3
APSE_K1B_STAGES=2 run r48_s2
