#!/bin/bash
# 8-GPU diagnosis of the round-1 SIGABRT (run under: gpurun --gpus 8 -- bash tools/diag_n8.sh)
#   A: the round-1 graph-capture mode ("global") with per-rank stderr kept  -> reproduces / names the failure
#   B: the default (thread-local capture)                                    -> the bench line at N=8
#   C: the on-hardware data-parallel correctness tests at world 2 and 8
O=gpurun_out/n8
mkdir -p $O
export TORCH_SHOW_CPP_STACKTRACES=1 NCCL_DEBUG=WARN
run() {  # name, timeout, env...
  name=$1; lim=$2; shift 2
  env "$@" ALDI_BENCH_DIAG_DIR=$O/$name timeout $lim python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > $O/$name.out 2> $O/$name.launcher.err
  echo "rc=$?" >> $O/$name.out
}
run global 150 ALDI_CAPTURE_MODE=global
run tl 240 ALDI_CAPTURE_MODE=thread_local
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/pytest.out 2>&1
echo "rc=$?" >> $O/pytest.out
tail -c 1500 $O/global.out; echo ----; tail -c 3000 $O/tl.out; echo ----; tail -20 $O/pytest.out
for f in $O/global/rank*.err; do echo "== $f"; grep -v "^$" $f | head -12; done 2>/dev/null | head -150
