set -x
timeout 600 python -m pytest tests/test_conv_umma_gpu.py tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu20.log 2>&1; tail -2 gpurun_out/pytest_gpu20.log
timeout 600 python tools/conv_bench.py > gpurun_out/cb_r1p.log 2>&1; grep -E "wgrad|total|top" gpurun_out/cb_r1p.log | head -50
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1p.log 2>&1; tail -1 gpurun_out/bench_r1p.log | cut -c1-200
