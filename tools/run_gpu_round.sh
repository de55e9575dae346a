set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu12.log 2>&1; tail -5 gpurun_out/pytest_gpu12.log
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench_epi.log 2>&1; grep -E "^G\.|^D1\.[0-3]|^E\.first|total|top" gpurun_out/conv_bench_epi.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1h.log 2>&1; tail -1 gpurun_out/bench_r1h.log | cut -c1-400
