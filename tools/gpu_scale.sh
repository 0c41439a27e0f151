#!/bin/bash
# strong-scaling series on one multi-GPU box: bench.py at N = $NS GPUs, workloads $WS; one JSON line each
mkdir -p gpurun_out
NS=${NS:-"1 2 4"}; WS=${WS:-"c5 c2"}; ARGS=${ARGS:-"--no-e2e --no-cpu-baseline --steps 8 --warmup 2"}
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_scale.py -q -k two_gpu -p no:cacheprovider 2>&1 | tail -3
for w in $WS; do for n in $NS; do
  if [ "$n" = "1" ]; then cmd="python bench.py"; else cmd="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py"; fi
  timeout 900 $cmd --gpus $n --workload $w $ARGS 2> gpurun_out/scale_${w}_$n.err | tail -1 > gpurun_out/scale_${w}_$n.json
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${w}_$n.json").read())
    print("$w N=$n value %.2f it/s ms %.3f exact_only %s hist/step %.3f ms allreduce %s frac %.3f" % (d["value"], d["ms_per_step"], d["exact_tier_only"] and round(d["exact_tier_only"]["value"],1), d["kernel_ms_per_step"].get("histogram",0), d["kernel_ms_per_step"].get("allreduce"), d["roofline"]["frac"]), d["kernel_ms_per_step"])
except Exception as e:
    print("$w N=$n FAILED", e); print(open("gpurun_out/scale_${w}_$n.err").read()[-1500:])
PY
done; done
