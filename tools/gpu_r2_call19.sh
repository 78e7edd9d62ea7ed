#!/bin/bash
# Round 2, GPU call 19 (1 GPU): surface-term kernel parity, the unmodified applications with Neumann terms, whole GPU suite
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== neumann parity"
timeout 900 python -m pytest tests/test_parity_gpu.py -q -m gpu -k "neumann" 2>&1 | tail -8
echo "== applications"
timeout 900 python -m pytest tests/test_zz_linear_constraints.py tests/test_reference_run.py -q -m gpu -k "mixed or neumann or application or reference_api" 2>&1 | tail -8
echo "== whole GPU suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8
} > $O/session19.log 2>&1
tail -40 $O/session19.log
