#!/bin/bash
# Round evidence, run on the GPU box (gpurun): launch list of a short bench run, `ncu --set full` of one
# steady-state default step and of the dense-bucket (10 M-sized buckets) step, summarised with
# tools/ncu_summary.py; the .ncu-rep files are dropped (gpurun brings back at most 64 MiB).
set -u
OUT=gpurun_out
TAG=${1:-r02}
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/${TAG}_launches_step.csv \
    python bench.py --steps 2 --warmup 1 --north-star-total 0 --no-representatives --no-cpu-baseline \
    > $OUT/${TAG}_launches_bench.log 2>&1
# step 3 of run_step.py = steady state (sizes learnt, one read-back); our kernels all end in "_kernel"
# (the first step has 32 launches that match, a steady-state step 37)
ncu --set full --clock-control none -k regex:"_kernel" -s 69 -c 37 -o /tmp/${TAG}_default \
    python tools/run_step.py --steps 3 > $OUT/${TAG}_prof_default.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_default.ncu-rep > $OUT/${TAG}_ncu_default.jsonl
if [ "${2:-}" != "nodense" ]; then
ncu --set full --clock-control none -k regex:"kmeans_t|ivf_assign|scan_tc|refine_block|vectorize_kernel" -s 40 -c 45 \
    -o /tmp/${TAG}_dense python tools/run_step.py --steps 2 --mass-range 1000 1280 > $OUT/${TAG}_prof_dense.log 2>&1
python tools/ncu_summary.py /tmp/${TAG}_dense.ncu-rep > $OUT/${TAG}_ncu_dense.jsonl
fi
wc -l $OUT/${TAG}_ncu_*.jsonl
