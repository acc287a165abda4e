#!/bin/bash
# PLAIN instantiation of the nv > 32 solver path: parity + bench lines of the two workloads it serves
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden_pipeline.py tests/test_gpu_api.py tests/test_gpu_parity.py -m gpu -q -k "g1 or three or sparse or G1 or replay" 2>&1 | tail -4
for wl in g1 three_humanoids; do
  timeout 400 python bench.py --workload $wl --steps 100 --warmup 20 > gpurun_out/r02_bench_$wl.json 2>> gpurun_out/r02_bench.err; echo "$wl rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$wl.json')); print(d['value'], d['e2e']['value'], d['kernel_ms'])"
done
