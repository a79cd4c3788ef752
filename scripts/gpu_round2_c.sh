#!/bin/bash
# round 2, run C: suite, fused/unfused A/B (new build), new-vs-previous build A/B on the unfused frame and the C3 step, sanitizers
mkdir -p gpurun_out
T=${TAG:-r02b}
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout=600 -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${T}_rc.txt
timeout -k 10 300 python scripts/ab_fused.py 3 10 > gpurun_out/${T}_ab_fused.log 2>&1; echo "ab rc=$?" | tee -a gpurun_out/${T}_rc.txt
echo "--- previous build (HEAD~) ---" >> gpurun_out/${T}_ab_fused.log
TEXPOSE_B200_LIB=$PWD/build/libtexpose_prev.so timeout -k 10 300 python scripts/ab_fused.py 2 10 >> gpurun_out/${T}_ab_fused.log 2>&1
for i in 1 2 3; do
  for lib in "" "$PWD/build/libtexpose_prev.so"; do
    echo -n "lib ${lib:-new}: " >> gpurun_out/${T}_ab_train.log; TEXPOSE_B200_LIB=$lib timeout -k 10 120 python scripts/step_timeline.py 20 2>&1 | grep "C3 train" >> gpurun_out/${T}_ab_train.log
  done
done
# sanitizers (smoke-sized; memcheck on everything, racecheck + synccheck on the render / train kernels and the peer kernels)
for tool in memcheck racecheck synccheck; do
  timeout -k 10 900 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py render train peer > gpurun_out/${T}_sanitizer_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${T}_rc.txt
  tail -3 gpurun_out/${T}_sanitizer_${tool}.log
done
cat gpurun_out/${T}_rc.txt; tail -4 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_ab_fused.log gpurun_out/${T}_ab_train.log
