#!/bin/bash
tag=${1:-r02q}
timeout 400 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/${tag}_tests.log; tail -3 gpurun_out/${tag}_tests.log
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 5 --warmup 2 --e2e-steps 1 --e2e-frames 120 --cpu-frames 0 --kernel-batches 1 > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/${tag}_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value']), round(d['ms_per_step'],2), round(d['pipeline_only']['ms_per_step'],2))" || tail -3 gpurun_out/${tag}_$name.err; }
run aux1 A=1
run aux0 APSE_SPARSE_AUX=0
