timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 400 python bench.py > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err; tail -c 400 gpurun_out/bench_e.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_k.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_q1hex_patch_affine -s 2 -c 1 -o gpurun_out/prof_patch_f python bench.py --steps 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_l.log 2>&1; tail -1 gpurun_out/ncu_l.log
