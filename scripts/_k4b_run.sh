{
python scripts/stamp_objective.py | grep "last CTA\|prologue\|result"
MG_OBJ_DEBUG=16 python scripts/stamp_objective.py | grep "last CTA\|prologue\|result"
for i in 1 2 3; do python scripts/time_objective.py; MG_OBJ_DEBUG=16 python scripts/time_objective.py; done
} > gpurun_out/r2b_timing.log 2>&1
cat gpurun_out/r2b_timing.log
