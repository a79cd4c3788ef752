# ncu source-level (SASS) stall samples of the training forward kernel; run under gpurun.
cd ${GRAFT_REPO_ROOT:-.}
timeout 250 ncu --set full --import-source on --clock-control none -k regex:${1:-nerf_stl_forward} -s 2 -c 1 -f -o /tmp/r01g_src python scripts/train_profile.py 3 > /tmp/cap_src.log 2>&1
ncu -i /tmp/r01g_src.ncu-rep --page source --csv > gpurun_out/r01g_${1:-nerf_stl_forward}_source.csv 2>/dev/null
wc -c gpurun_out/r01g_${1:-nerf_stl_forward}_source.csv
