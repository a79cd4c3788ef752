#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r02c}
for lib in "" build/libtexpose_prev.so; do echo "== lib ${lib:-new}" >> gpurun_out/${T}_fused_diff.log; TEXPOSE_B200_LIB=${lib:+$PWD/$lib} timeout -k 10 200 python scripts/fused_diff.py >> gpurun_out/${T}_fused_diff.log 2>&1; done
for r in 1 2; do for lib in "" build/libtexpose_exp1.so build/libtexpose_exp2.so build/libtexpose_exp4.so build/libtexpose_exp7.so; do
  echo "== lib ${lib:-new}" >> gpurun_out/${T}_ab_exp.log; TEXPOSE_B200_LIB=${lib:+$PWD/$lib} timeout -k 10 200 python scripts/ab_fused.py 1 10 >> gpurun_out/${T}_ab_exp.log 2>&1; done; done
timeout -k 10 600 compute-sanitizer --tool synccheck --print-limit 400 python scripts/sanitize_target.py render > gpurun_out/${T}_sanitizer_synccheck_render.log 2>&1
grep "shared address\|in mlp_tc.cu" gpurun_out/${T}_sanitizer_synccheck_render.log | sort | uniq -c | sort -rn | head -20
cat gpurun_out/${T}_fused_diff.log gpurun_out/${T}_ab_exp.log
