#!/bin/bash
# ncu --set full captures (one launch each) of the kernels the round's numbers rest on; raw pages are summarised into profiles/ by hand
mkdir -p gpurun_out
T=${TAG:-r02full}
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:nerf_stl_forward -c 1 -o gpurun_out/${T}_render -f \
    python bench.py --steps 1 --warmup 0 --no-train --no-cpu-baseline --no-parity-frame > /dev/null 2>&1; echo "render rc=$?"
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:"backward_chain_fused|dw_gemm_kernel|nerf_stl_forward" -s 6 -c 3 -o gpurun_out/${T}_train -f \
    python scripts/train_profile.py 3 > /dev/null 2>&1; echo "train rc=$?"
timeout -k 10 600 ncu --set full --import-source on --clock-control none -k regex:"chain_backward_staged|nerf_forward_split" -s 4 -c 2 -o gpurun_out/${T}_plain -f \
    python scripts/plain_train_time.py > /dev/null 2>&1; echo "plain rc=$?"
for n in render train plain; do ncu -i gpurun_out/${T}_$n.ncu-rep --page raw --csv > gpurun_out/${T}_${n}_raw.csv 2>/dev/null; done
ls -la gpurun_out/${T}_*
