set -x
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r1m_2gpu.log 2>&1; echo rc=$?; grep -E '^\{' gpurun_out/bench_r1m_2gpu.log | cut -c1-400
timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 2 --warmup 3 --impl reference > gpurun_out/bench_ref_2gpu.log 2>&1; echo rc=$?; grep -E '^\{' gpurun_out/bench_ref_2gpu.log | cut -c1-300
