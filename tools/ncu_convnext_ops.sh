#!/bin/bash
# one ncu --set full capture of the LayerNorm / layer-scale stream kernels at the ConvNeXt-L stage-0 geometry
timeout 70 ncu --set full --clock-control none --import-source on -k "regex:ln_fwd_vec_kernel|ln_bwd_vec_kernel|layerscale_fwd_kernel" \
  --launch-skip 4 -c 3 -f -o gpurun_out/s28_ln python tools/bench_convnext_ops.py --iters 1 \
  --only layernorm_forward,layernorm_backward,layerscale_forward --stages 0 > gpurun_out/s28_ncu.log 2>&1
tail -4 gpurun_out/s28_ncu.log
