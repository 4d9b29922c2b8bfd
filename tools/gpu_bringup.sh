#!/bin/bash
# First-contact GPU run: every step in its own process (a trapped kernel poisons
# the CUDA context), each under its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
L=gpurun_out/bringup.log
: > $L
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv >> $L 2>&1
run() { echo "=== $*" >> $L; timeout "$@" >> $L 2>&1; echo "=== exit $?" >> $L; }
run 600 python -m pytest tests/test_gpu_stages.py -x -q -p no:cacheprovider --timeout 300 \
    -k "hash_table or vectorize or bucket_sort or dbscan or split or rt_tolerance"
run 600 python -m pytest tests/test_gpu_stages.py -q -p no:cacheprovider --timeout 300 \
    -k "knn_csr_exhaustive_exact and 1]"
run 300 python -m pytest tests/test_gpu_stages.py -x -q -p no:cacheprovider --timeout 200 -k "scan_pairs"
run 600 python -m pytest tests/test_gpu_stages.py -q -p no:cacheprovider --timeout 300 -k "knn_csr or ivf or kmeans"
run 900 python -m pytest tests/test_gpu_pipeline.py -q -p no:cacheprovider --timeout 600 -k "not full_size"
tail -5 $L
