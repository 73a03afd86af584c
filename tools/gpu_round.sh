#!/bin/bash
# final visit: every GPU test, smoke, the default bench line
tag=${1:-r02o}
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/${tag}_gputest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${tag}_gputest.log; tail -3 gpurun_out/${tag}_gputest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cut -c1-200 gpurun_out/${tag}_bench_n1.json
