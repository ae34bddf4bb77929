#!/bin/bash
# One GPU box, end of a change: smoke, the GPU test-suite, the bench line, a `--set full` capture of K4b and the launch list of a short
# bench run (each ncu pass only after the plain run).    gpurun --timeout 1200 -- 'bash scripts/gpu_round_check.sh'
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
(time python -m pytest tests -m gpu -x -q) > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err
ncu --set full --clock-control none --import-source on -k regex:objective_stream -s 2 -c 1 -o gpurun_out/k4b -f python scripts/profile_objective.py > gpurun_out/ncu_k4b.log 2>&1; tail -2 gpurun_out/ncu_k4b.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
