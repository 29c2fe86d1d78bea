#!/bin/bash
# the driver's own 8-GPU command on the final tree (supervised bench), strict time limit
O=gpurun_out/n8final; mkdir -p $O
timeout 130 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29591 \
  bench.py --gpus 8 --steps 20 --warmup 5 > $O/bench.out 2> $O/bench.err; echo "rc=$?" | tee $O/rc.txt
grep "^{" $O/bench.out | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('value','ms_per_step','attempts','restarts','n_gpus')}, 'e2e', d['e2e']['value'])"
tail -c 600 $O/bench.err
