#!/bin/bash
# multi-GPU validation on one box: the 2-GPU parity test, then the default bench line (C2 + extra C5) at N = $NS ranks
mkdir -p gpurun_out
NS=${NS:-"1 2 4"}
nvidia-smi -L | head -8 > gpurun_out/multi_gpus.txt
timeout 900 python -m pytest tests/test_gpu_scale.py -q -k two_gpu -p no:cacheprovider > gpurun_out/pytest_2gpu.txt 2>&1; echo "2-GPU test rc=$?"; tail -3 gpurun_out/pytest_2gpu.txt
for n in $NS; do
  if [ "$n" = "1" ]; then launcher="python"; else launcher="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2960$n"; fi
  timeout 900 $launcher bench.py --gpus $n --steps 20 --warmup 5 ${BENCH_ARGS} \
      > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "bench N=$n rc=$?"
  grep '^{' gpurun_out/bench_n$n.json | cut -c1-400
done
