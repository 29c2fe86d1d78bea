#!/bin/bash
# round-end verification: the full GPU suite, the smoke entry, the default bench line
(timeout 280 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/sNN_tests.log
cat gpurun_out/sNN_tests.log
timeout 90 python __graft_entry__.py smoke > gpurun_out/sNN_smoke.log 2>&1; tail -3 gpurun_out/sNN_smoke.log
timeout 200 python bench.py > gpurun_out/sNN_bench.json 2> gpurun_out/sNN_bench.err; cat gpurun_out/sNN_bench.json; tail -3 gpurun_out/sNN_bench.err
