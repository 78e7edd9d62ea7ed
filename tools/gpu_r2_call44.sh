#!/bin/bash
# Round 2, GPU call 44 (1 GPU, the last seconds of the budget): sampled conductivity through the atomic-free path as the default
mkdir -p gpurun_out/r2
{
timeout 40 python -m pytest tests/test_parity_gpu.py tests/test_reference_run.py -q -m gpu -x -k "kappafun" 2>&1 | tail -3
} > gpurun_out/r2/session44.log 2>&1
cat gpurun_out/r2/session44.log
