#!/bin/bash
# Runs on the B200 box (via gpurun): ncu evidence for profiles/.  Numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
# (1) launch list of the bench command (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG:-r01}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/${TAG:-r01}_bench_under_ncu.log 2>&1
# (2) the dominant kernel at full C2 size: DRAM traffic + tensor pipe
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum \
    --clock-control none -k regex:nerf_stl_forward -c 1 --csv --log-file gpurun_out/${TAG:-r01}_tc_full_metrics.csv \
    python bench.py --steps 1 --warmup 0 --no-train --no-cpu-baseline > /dev/null 2>&1
# (3) HBM-bound kernels at C2 size
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"composite|raygen|sample_depth|box_range|ray_bias|gather_rows" -c 12 --csv --log-file gpurun_out/${TAG:-r01}_hbm_kernels.csv \
    python bench.py --steps 1 --warmup 1 --no-train --no-cpu-baseline > /dev/null 2>&1
# (4) clocks line during a plain run
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG:-r01}_clocks.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/${TAG:-r01}_bench.json 2> gpurun_out/${TAG:-r01}_bench.err
kill $SMI
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG:-r01}_bench_reference.json 2>> gpurun_out/${TAG:-r01}_bench.err
tail -c 3000 gpurun_out/${TAG:-r01}_bench.json
# (5) training step: launch list + DRAM traffic of the tensor-core backward kernels
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG:-r01}_train_launches.csv python scripts/train_profile.py 3 > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"backward_chain|dw_gemm|nerf_stl_forward|bwd_finish|dw_reduce|patch_loss|composite_stl" -s 9 -c 18 --csv --log-file gpurun_out/${TAG:-r01}_train_kernels.csv \
    python scripts/train_profile.py 2 > /dev/null 2>&1
