#!/bin/bash
O=gpurun_out/n2sup; mkdir -p $O
run() { name=$1; shift; env "$@" timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
  --master-port 29581 bench.py --gpus 2 --steps 20 --warmup 5 > $O/$name.out 2> $O/$name.err; echo "$name rc=$?" | tee -a $O/summary.txt; }
run plain A=1
run inject ALDI_BENCH_INJECT_FAIL=1:0
for n in plain inject; do python - <<PY
import json
try:
    l=[x for x in open("$O/$n.out") if x.startswith("{")][-1]; d=json.loads(l)
    print("$n", {k:d.get(k) for k in ("value","ms_per_step","attempts","restarts","n_gpus")}, "e2e", d["e2e"]["value"])
except Exception as e: print("$n: no line", e)
PY
done
tail -4 $O/inject.err | cut -c1-300
