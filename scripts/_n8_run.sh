mkdir -p gpurun_out
N=${N:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/h2d_bandwidth.py > gpurun_out/n${N}_h2d.log 2>&1; tail -3 gpurun_out/n${N}_h2d.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; tail -3 gpurun_out/n${N}_bench.err; cut -c1-400 gpurun_out/n${N}_bench.json
