set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu31.log 2>&1; tail -3 gpurun_out/pytest_gpu31.log
timeout 300 python -m pytest tests/test_step_gpu.py -x -q -m gpu -s -k "fp32_engine" 2>&1 | grep -E "losses" | cut -c1-250
SRGAN_DBG_NO_BATCH_FAKES=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-140 | sed 's/^/SEP /'
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu 2>&1 | tail -1 | cut -c1-140 | sed 's/^/BATCHED /'
SRGAN_DBG_NO_BATCH_FAKES=1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --batch 8 2>&1 | tail -1 | cut -c1-140 | sed 's/^/SEP8 /'
timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu --batch 8 2>&1 | tail -1 | cut -c1-140 | sed 's/^/BATCHED8 /'
