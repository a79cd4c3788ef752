#!/bin/bash
# evidence for the split-fp16 kernel only: ncu metrics of one launch, ncu --set full capture, the three sanitizers
mkdir -p gpurun_out
T=${TAG:-r02s}
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum \
    --clock-control none -k regex:nerf_forward_split -s 2 -c 1 --csv --log-file gpurun_out/${T}_split_kernel_metrics.csv \
    env TP_SPLIT_ONLY=1 python scripts/fp32_frame.py > /dev/null 2>&1; echo "ncu split rc=$?" | tee gpurun_out/${T}_rc.txt
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:nerf_forward_split -s 2 -c 1 -o gpurun_out/${T}_split_full -f \
    env TP_SPLIT_ONLY=1 python scripts/fp32_frame.py > /dev/null 2>&1; echo "ncu full rc=$?" | tee -a gpurun_out/${T}_rc.txt
ncu -i gpurun_out/${T}_split_full.ncu-rep --page raw --csv > gpurun_out/${T}_split_full_raw.csv 2>/dev/null
for tool in memcheck racecheck synccheck; do
  timeout -k 10 600 compute-sanitizer --tool $tool --print-limit 30 python scripts/sanitize_target.py split > gpurun_out/${T}_sanitizer_split_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${T}_rc.txt
done
cat gpurun_out/${T}_rc.txt; cat gpurun_out/${T}_split_kernel_metrics.csv | tail -8; for tool in memcheck racecheck synccheck; do tail -3 gpurun_out/${T}_sanitizer_split_${tool}.log; done
