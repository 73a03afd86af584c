#!/bin/bash
tag=${1:-r02g}
timeout 300 python tools/profile_postpass.py > gpurun_out/${tag}_postpass.log 2>&1; tail -4 gpurun_out/${tag}_postpass.log
bash tools/gpu_profile.sh $tag
