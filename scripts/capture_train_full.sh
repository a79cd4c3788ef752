set -x
cd $GRAFT_REPO_ROOT
for k in nerf_stl_forward backward_chain_fused; do
  timeout 250 ncu --set full --clock-control none -k regex:$k -s 2 -c 1 -f -o /tmp/r01g_$k python scripts/train_profile.py 3 > /tmp/cap_$k.log 2>&1
  ncu -i /tmp/r01g_$k.ncu-rep --page raw --csv > gpurun_out/r01g_train_${k}_full_raw.csv 2>/dev/null
  ls -la /tmp/r01g_$k.ncu-rep
done
wc -c gpurun_out/r01g_train_*_full_raw.csv
