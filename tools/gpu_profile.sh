#!/bin/bash
# profile visit (1 GPU): final bench lines (default, config 1, config 4), launch list and ncu full of the preprocess kernels.
# Every ncu command runs only after the same program exited 0 without ncu.  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02g}
timeout 600 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; cut -c1-300 gpurun_out/${tag}_bench_n1.json
timeout 300 python bench.py --workload preprocess64 > gpurun_out/${tag}_bench_config1.json 2> gpurun_out/${tag}_bench_config1.err; cut -c1-200 gpurun_out/${tag}_bench_config1.json
timeout 400 python bench.py --workload dense-apriltag > gpurun_out/${tag}_bench_config4_apriltag.json 2> gpurun_out/${tag}_bench_config4_apriltag.err; cut -c1-200 gpurun_out/${tag}_bench_config4_apriltag.json
timeout 400 python bench.py --workload dense-classic > gpurun_out/${tag}_bench_config4_classic.json 2> gpurun_out/${tag}_bench_config4_classic.err; cut -c1-200 gpurun_out/${tag}_bench_config4_classic.json
B=20 STEPS=2 timeout 300 python tools/ncu_target.py > gpurun_out/${tag}_target.log 2>&1 || exit 1
B=20 STEPS=2 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 400 --csv \
  --log-file gpurun_out/${tag}_launches.csv python tools/ncu_target.py > gpurun_out/${tag}_ncu1.log 2>&1
B=20 STEPS=2 timeout 900 ncu --set full --clock-control none --import-source on -k "regex:k_preprocess_tma|k_sparse" -s 3 -c 3 \
  -o gpurun_out/prof_${tag}_sparse -f python tools/ncu_target.py > gpurun_out/${tag}_ncu2.log 2>&1
ls -la gpurun_out/ | tail -20
