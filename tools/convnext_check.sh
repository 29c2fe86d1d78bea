#!/bin/bash
# ConvNeXt stream kernels: op + model + step tests, per-stage micro-benchmark, ConvNeXt-L step bench with profile
(timeout 400 python -m pytest tests/test_gpu_convnext_ops.py tests/test_gpu_convnext.py tests/test_gpu_step_parity.py -k "convnext or layernorm or dwconv or gelu" -x -q 2>&1 | tail -15) > gpurun_out/sNN_tests.log

timeout 150 python tools/bench_convnext.py --size L --ims 2 --steps 3 --profile > gpurun_out/sNN_convnext.json 2> gpurun_out/sNN_convnext.err
cat gpurun_out/sNN_tests.log gpurun_out/sNN_convnext.json; head -12 gpurun_out/sNN_convnext.err
