#!/bin/bash
tag=${1:-r02h}
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_pipeline.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_tests.log
[ $rc -ne 0 ] && exit $rc
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 6 --warmup 3 --e2e-steps 1 --cpu-frames 0 $EXTRA > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/${tag}_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value']), round(d['ms_per_step'],2), round(d['pipeline_only']['ms_per_step'],2), d['roofline']['kernel'], round(d['roofline']['frac'],4), round(d['roofline']['avg_ms'],4), d['postpass_ms_per_step'])"; }
run nreg40 A=1
run nreg48 APSE_K1B_NREG=48
run nreg40_b A=1
run nreg48_b APSE_K1B_NREG=48
