set -x
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/bench_r1w_8gpu.log 2>&1; echo rc=$?; grep -E '^\{' gpurun_out/bench_r1w_8gpu.log | cut -c1-300
