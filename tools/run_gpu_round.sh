set -x
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 4 --steps 8 --warmup 3 > gpurun_out/bench_r1x_4gpu.log 2>&1; echo rc=$?; grep -E '^\{' gpurun_out/bench_r1x_4gpu.log | cut -c1-200
