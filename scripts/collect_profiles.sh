#!/bin/bash
# Runs on the B200 box (via gpurun): the round's evidence for profiles/.  Numbers printed under ncu are never bench values.
#   TAG=r02 bash scripts/collect_profiles.sh        -> gpurun_out/${TAG}_*  (copy what should be judged into profiles/)
mkdir -p gpurun_out
T=${TAG:-r02}
# (1) the whole GPU suite
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout=900 -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${T}_rc.txt
# (1b) the fused render launch at full C2 size: DRAM traffic + tensor pipe, keyed by the source / .so hash -- BEFORE the plain bench,
# whose JSON line then carries roofline.traffic of this very build
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum \
    --clock-control none -k regex:nerf_stl_forward -c 1 --csv --log-file gpurun_out/${T}_render_full_c2_metrics.csv \
    python bench.py --steps 1 --warmup 0 --no-train --no-cpu-baseline > /dev/null 2>&1; echo "ncu render rc=$?" | tee -a gpurun_out/${T}_rc.txt
python - <<PY
import csv, hashlib, json
rows = [r for r in csv.reader(open("gpurun_out/${T}_render_full_c2_metrics.csv")) if len(r) > 5]
hdr = rows[0]; mi, vi = hdr.index("Metric Name"), hdr.index("Metric Value")
m = {r[mi]: float(r[vi].replace(",", "")) for r in rows[1:]}
sha = hashlib.sha256(open("texpose_b200/libtexpose_b200.so", "rb").read()).hexdigest()[:16]
import sys; sys.path.insert(0, "."); import bench
out = dict(so_sha16=sha, src_sha16=bench.src_sha16(), dram_bytes_per_launch=int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]), dram_read=int(m["dram__bytes_read.sum"]),
           dram_write=int(m["dram__bytes_write.sum"]), kernel_ns=m.get("gpu__time_duration.sum"),
           tensor_pipe_active_pct=m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"), l2_bytes=m.get("lts__t_bytes.sum"),
           source="ncu --clock-control none, one launch of the fused render kernel at C2 size (480x640x128), units as printed by ncu")
json.dump(out, open("gpurun_out/${T}_render_traffic.json", "w"), indent=1)
json.dump(out, open("profiles/r02_render_traffic.json", "w"), indent=1)
print(out)
PY
# (2) plain bench runs (ours, then the reference arm) with a clocks log beside them
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${T}_clocks_during_bench.csv &
SMI=$!
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${T}_rc.txt
kill $SMI
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_n1.json 2>> gpurun_out/${T}_bench.err; echo "ref rc=$?" | tee -a gpurun_out/${T}_rc.txt
# (3) launch list of the bench command (cold-cache, serialised: compare SHARES)
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "ncu launches rc=$?" | tee -a gpurun_out/${T}_rc.txt
# (4b) the split-fp16 kernel of the fp32-parity mode: one launch (2^21 samples of the C2 frame), tensor pipe + traffic
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum \
    --clock-control none -k regex:nerf_forward_split -s 2 -c 1 --csv --log-file gpurun_out/${T}_split_kernel_metrics.csv \
    env TP_SPLIT_ONLY=1 python scripts/fp32_frame.py > /dev/null 2>&1; echo "ncu split rc=$?" | tee -a gpurun_out/${T}_rc.txt
# (5) training step: launch list + per-kernel DRAM traffic / tensor pipe
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${T}_launches_train_step.csv python scripts/train_profile.py 3 > /dev/null 2>&1
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"backward_chain|dw_gemm|nerf_stl_forward|bwd_finish|dw_reduce|patch_loss|composite_stl" -s 9 -c 18 --csv --log-file gpurun_out/${T}_train_kernels_metrics.csv \
    python scripts/train_profile.py 2 > /dev/null 2>&1
# (6) sanitizers on smoke-sized launches of every hand-written synchronisation protocol (peer kernels: scripts/gpu_round2_n2.sh, two GPUs)
for tool in memcheck racecheck synccheck; do
  for what in render train split plain; do      # one process per target: a tool that aborts one target does not hide the others
    timeout -k 10 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_target.py $what > gpurun_out/${T}_sanitizer_${tool}_${what}.log 2>&1
    echo "$tool $what rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${T}_sanitizer_${tool}_${what}.log | head -1)" | tee -a gpurun_out/${T}_rc.txt
  done
done
cat gpurun_out/${T}_rc.txt; tail -4 gpurun_out/${T}_pytest.log; head -c 700 gpurun_out/${T}_bench_n1.json
