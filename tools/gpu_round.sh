#!/bin/bash
tag=${1:-r02l}
timeout 400 python bench.py --gpus 8 --e2e-steps 1 --cpu-frames 0 > gpurun_out/${tag}_n8.json 2> gpurun_out/${tag}_n8.err; cut -c1-200 gpurun_out/${tag}_n8.json
timeout 400 python bench.py --gpus 4 --e2e-steps 1 --cpu-frames 0 --steps 6 > gpurun_out/${tag}_n4.json 2> gpurun_out/${tag}_n4.err; cut -c1-200 gpurun_out/${tag}_n4.json
timeout 400 python bench.py --gpus 2 --e2e-steps 1 --cpu-frames 0 --steps 6 > gpurun_out/${tag}_n2.json 2> gpurun_out/${tag}_n2.err; cut -c1-200 gpurun_out/${tag}_n2.json
