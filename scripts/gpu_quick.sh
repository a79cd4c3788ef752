#!/bin/bash
# quick check of a render-kernel change: view-bias rows vs table, fused vs multi-kernel outputs, fused tests, same-box A/B
mkdir -p gpurun_out
T=${TAG:-quick}
timeout -k 10 120 python scripts/view_bias_debug.py > gpurun_out/${T}_view_bias_debug.log 2>&1
timeout -k 10 200 python scripts/fused_diff.py > gpurun_out/${T}_fused_diff.log 2>&1
timeout -k 10 600 python -m pytest tests/test_gpu_fused_render.py tests/test_gpu_tc.py -q --timeout=600 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${T}_rc.txt
for r in 1 2 3; do timeout -k 10 200 python scripts/ab_fused.py 1 10 >> gpurun_out/${T}_ab_fused.log 2>&1; done
cat gpurun_out/${T}_view_bias_debug.log gpurun_out/${T}_fused_diff.log; tail -5 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_ab_fused.log
