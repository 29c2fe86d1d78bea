#!/bin/bash
# final-settings check at 8 GPUs: the driver's own command (supervised), restart path with an injected failure, DP tests
O=gpurun_out/n8c; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
  --master-port 29571 bench.py --gpus 8 --steps 20 --warmup 5 > $O/$name.out 2> $O/$name.err; echo "$name rc=$?" | tee -a $O/summary.txt; }
run plain A=1
run inject ALDI_BENCH_INJECT_FAIL=3:0
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest.out 2>&1; echo "pytest rc=$?" | tee -a $O/summary.txt
for n in plain inject; do python - <<PY
import json
try:
    l=[x for x in open("$O/$n.out") if x.startswith("{")][-1]; d=json.loads(l)
    print("$n", {k:d.get(k) for k in ("value","ms_per_step","attempts","restarts","n_gpus")}, "e2e", d["e2e"]["value"])
except Exception as e: print("$n: no line", e)
PY
done
tail -5 $O/pytest.out; tail -5 $O/inject.err | cut -c1-300
