#!/bin/bash
tag=${1:-r02j}
timeout 900 python -m pytest tests/test_gpu_sparse.py tests/test_gpu_pipeline.py tests/test_gpu_preprocess.py -x -q -m gpu > gpurun_out/${tag}_tests.log 2>&1
rc=$?; echo "pytest rc=$rc" >> gpurun_out/${tag}_tests.log; tail -4 gpurun_out/${tag}_tests.log
[ $rc -ne 0 ] && exit $rc
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 6 --warmup 3 --e2e-steps 1 --cpu-frames 0 $EXTRA > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/${tag}_$name.json').read().strip().splitlines()[-1]); k=d['kernel_ms_per_frame']; print('$name', round(d['value']), round(d['ms_per_step'],2), round(d['pipeline_only']['ms_per_step'],2), d['roofline']['kernel'], round(d['roofline']['frac'],4), 'exact us/frame', round(1e3*k['k_sparse_exact'],2), 'bounds', round(1e3*k['k_preprocess_fused'],2))"; }
run ctas4 A=1
run ctas6 APSE_EXACT_CTAS=6
run ctas8 APSE_EXACT_CTAS=8
run ctas3 APSE_EXACT_CTAS=3
timeout 300 python bench.py --workload preprocess64 2>/dev/null | cut -c1-160
