#!/bin/bash
# first validation of the round-2 kernels: new tests first (own timeout), then the whole suite, then a short bench
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2 > gpurun_out/a_gpu.txt
timeout -k 10 600 python -m pytest tests/test_gpu_fused_render.py -q --timeout=300 -s > gpurun_out/a_fused.log 2>&1; echo "fused rc=$?" | tee -a gpurun_out/a_rc.txt
timeout -k 10 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_dropin.py tests/test_gpu_parity_c3.py -q --timeout=600 -s > gpurun_out/a_parity.log 2>&1; echo "parity rc=$?" | tee -a gpurun_out/a_rc.txt
timeout -k 10 900 python -m pytest tests -m gpu -q --timeout=300 --deselect tests/test_gpu_fused_render.py --deselect tests/test_gpu_parity_c3.py > gpurun_out/a_suite.log 2>&1; echo "suite rc=$?" | tee -a gpurun_out/a_rc.txt
timeout -k 10 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" | tee -a gpurun_out/a_rc.txt
tail -3 gpurun_out/a_fused.log; tail -3 gpurun_out/a_parity.log; tail -3 gpurun_out/a_suite.log; head -c 1500 gpurun_out/a_bench.json
