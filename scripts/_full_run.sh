mkdir -p gpurun_out
(time python -m pytest tests -m gpu -x -q) > gpurun_out/f_pytest.log 2>&1; tail -4 gpurun_out/f_pytest.log
python bench.py > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err; tail -2 gpurun_out/f_bench.err
ncu --set full --clock-control none --import-source on -k regex:objective_stream -s 2 -c 1 -o gpurun_out/r2e_k4b -f python scripts/profile_objective.py > gpurun_out/f_ncu.log 2>&1; tail -2 gpurun_out/f_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches.csv python bench.py --steps 8 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/f_ncu2.log 2>&1; tail -1 gpurun_out/f_ncu2.log | cut -c1-200
