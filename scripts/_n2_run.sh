mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; tail -3 gpurun_out/n${N}_bench.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/n${N}_ref.json 2> gpurun_out/n${N}_ref.err; tail -2 gpurun_out/n${N}_ref.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 scripts/h2d_bandwidth.py > gpurun_out/n${N}_h2d.log 2>&1; tail -12 gpurun_out/n${N}_h2d.log
python scripts/time_host_overhead.py > gpurun_out/host_overhead.log 2>&1; tail -15 gpurun_out/host_overhead.log
