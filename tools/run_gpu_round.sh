set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu23.log 2>&1; tail -2 gpurun_out/pytest_gpu23.log
SRGAN_DBG_NORM_CTAS_PER_SM=8 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-140 | sed 's/^/C8 /'
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-140 | sed 's/^/C4 /'
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --batch 32 2>&1 | tail -1 | cut -c1-140 | sed 's/^/C4b32 /'
