#!/bin/bash
# two GPUs: strong-scaling bench (one frame, two row blocks), frame-gather / peer tests over real NVLink, sanitizers on the peer kernels
mkdir -p gpurun_out
T=${TAG:-r02n2}
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/${T}_gpus.txt
timeout -k 10 600 python -m pytest tests/test_gpu_frame_gather.py tests/test_gpu_peer.py -q --timeout=400 > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee gpurun_out/${T}_rc.txt
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "bench2 rc=$?" | tee -a gpurun_out/${T}_rc.txt
for tool in memcheck racecheck; do
  timeout -k 10 600 compute-sanitizer --tool $tool --target-processes all --print-limit 20 python -m pytest tests/test_gpu_frame_gather.py::test_frame_gather_equals_single_gpu_frame[2] \
      tests/test_gpu_peer.py::test_two_processes_through_cuda_ipc -q --timeout=500 > gpurun_out/${T}_sanitizer_peer_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/${T}_rc.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed" gpurun_out/${T}_sanitizer_peer_${tool}.log | tail -6
done
cat gpurun_out/${T}_rc.txt; tail -3 gpurun_out/${T}_pytest.log; head -c 1200 gpurun_out/${T}_bench_n2.json; tail -5 gpurun_out/${T}_bench_n2.err
