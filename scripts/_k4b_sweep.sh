{
python -m pytest tests -m gpu -x -q -k "objective or Objective or fused or acoustic or reuse or random_walk" 2>&1 | tail -3
python scripts/stamp_objective.py 2>&1 | tail -8
python scripts/time_objective.py
python scripts/time_objective.py
B=1024 python scripts/time_objective.py
} > gpurun_out/k4b_sweep.log 2>&1
cat gpurun_out/k4b_sweep.log
