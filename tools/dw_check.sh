#!/bin/bash
# depthwise-kernel check: op tests, ConvNeXt model tests, ConvNeXt-L step bench (new vs legacy kernels)
(timeout 300 python -m pytest tests/test_gpu_convnext_ops.py tests/test_gpu_convnext.py -x -q 2>&1 | tail -15) > gpurun_out/s21_tests.log
timeout 120 python tools/bench_convnext.py --size L --ims 2 --steps 3 > gpurun_out/s21_convnext.json 2> gpurun_out/s21_convnext.err
ALDI_DW7_LEGACY=1 timeout 120 python tools/bench_convnext.py --size L --ims 2 --steps 3 > gpurun_out/s21_convnext_legacy.json 2>> gpurun_out/s21_convnext.err
cat gpurun_out/s21_tests.log gpurun_out/s21_convnext.json gpurun_out/s21_convnext_legacy.json; tail -5 gpurun_out/s21_convnext.err
