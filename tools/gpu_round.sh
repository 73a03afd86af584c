#!/bin/bash
# one GPU-box visit: variants of the streamed post-pass at N=1, three repeats each (outputs under gpurun_out/<tag>_*)
tag=${1:-r02e}
run() { name=$1; shift; env "$@" timeout 600 python bench.py --steps 6 --warmup 3 --e2e-steps 1 --cpu-frames 0 --kernel-batches 1 $EXTRA > gpurun_out/${tag}_$name.json 2> gpurun_out/${tag}_$name.err; python -c "
import json,sys; d=json.loads(open('gpurun_out/${tag}_$name.json').read().strip().splitlines()[-1]); print('$name', round(d['value']), round(d['ms_per_step'],2), round(d['pipeline_only']['ms_per_step'],2))"; }
for rep in 1 2 3; do
EXTRA=--no-stream run nostream_$rep A=1
EXTRA= run inline_$rep APSE_POST_THREAD=0
run thread_sync_$rep APSE_POST_THREAD=1 APSE_POST_POLL=0
run thread_sync_sw1e4_$rep APSE_POST_THREAD=1 APSE_POST_POLL=0 APSE_POST_SWITCH=0.0001
run thread_poll_sw5e4_$rep APSE_POST_THREAD=1 APSE_POST_POLL=1 APSE_POST_SWITCH=0.0005
done
nproc; cat /proc/cpuinfo | grep MHz | sort | uniq -c | head -5
