#!/bin/bash
# Round 2, GPU call 20 (1 GPU): convection kernel parity, whole GPU suite, smoke, default bench + launch list
mkdir -p gpurun_out/r2
O=gpurun_out/r2
export PYTHONUNBUFFERED=1
{
echo "== convection parity"
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_reference_run.py -q -m gpu -k "convection" 2>&1 | tail -8
echo "== whole GPU suite"
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -8
echo "== smoke"
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -3
} > $O/session20.log 2>&1
tail -30 $O/session20.log
