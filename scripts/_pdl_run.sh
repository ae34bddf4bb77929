{
python -m pytest tests -m gpu -x -q -k "upsample or objective or fused or scan or packed" 2>&1 | tail -3
for i in 1 2; do
MG_PDL=1 python bench.py --no-extras --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL=1', d['ms_per_step'], d['value'], d['roofline']['avg_launch_ms'], d['roofline_k4b']['avg_launch_ms'])"
MG_PDL=0 python bench.py --no-extras --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('PDL=0', d['ms_per_step'], d['value'], d['roofline']['avg_launch_ms'], d['roofline_k4b']['avg_launch_ms'])"
done
} > gpurun_out/pdl.log 2>&1
cat gpurun_out/pdl.log
