#!/bin/bash
# round-2 validation + first measurements: whole GPU suite, fused/unfused A/B, bench (N=1), reference arm, ncu launch list and
# DRAM traffic of the fused render launch (keyed by the .so hash)
mkdir -p gpurun_out
T=${TAG:-r02a}
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout=600 -s > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${T}_rc.txt
timeout -k 10 300 python scripts/ab_fused.py 3 10 > gpurun_out/${T}_ab_fused.log 2>&1; echo "ab rc=$?" | tee -a gpurun_out/${T}_rc.txt
timeout -k 10 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" | tee -a gpurun_out/${T}_rc.txt
timeout -k 10 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_reference_n1.json 2>> gpurun_out/${T}_bench.err; echo "ref rc=$?" | tee -a gpurun_out/${T}_rc.txt
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-train --no-cpu-baseline > gpurun_out/${T}_bench_under_ncu.log 2>&1; echo "ncu1 rc=$?" | tee -a gpurun_out/${T}_rc.txt
timeout -k 10 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.avg.per_second,lts__t_bytes.sum \
    --clock-control none -k regex:nerf_stl_forward -c 1 --csv --log-file gpurun_out/${T}_render_full_c2_metrics.csv \
    python bench.py --steps 1 --warmup 0 --no-train --no-cpu-baseline > /dev/null 2>&1; echo "ncu2 rc=$?" | tee -a gpurun_out/${T}_rc.txt
python - <<PY
import csv, hashlib, json
rows = [r for r in csv.reader(open("gpurun_out/${T}_render_full_c2_metrics.csv")) if len(r) > 5]
hdr = rows[0]; mi, vi = hdr.index("Metric Name"), hdr.index("Metric Value")
m = {r[mi]: float(r[vi].replace(",", "")) for r in rows[1:]}
sha = hashlib.sha256(open("texpose_b200/libtexpose_b200.so", "rb").read()).hexdigest()[:16]
out = dict(so_sha16=sha, dram_bytes_per_launch=int(m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]), dram_read=int(m["dram__bytes_read.sum"]),
           dram_write=int(m["dram__bytes_write.sum"]), kernel_ns=m.get("gpu__time_duration.sum"),
           tensor_pipe_active_pct=m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"), l2_bytes=m.get("lts__t_bytes.sum"),
           source="ncu --clock-control none, one launch of the fused render kernel at C2 size (480x640x128), units as printed by ncu")
json.dump(out, open("gpurun_out/${T}_render_traffic.json", "w"), indent=1)
print(out)
PY
cat gpurun_out/${T}_rc.txt; tail -4 gpurun_out/${T}_pytest.log; cat gpurun_out/${T}_ab_fused.log | tail -3; head -c 900 gpurun_out/${T}_bench_n1.json
