set -x
timeout 600 python -m pytest tests/test_step_gpu.py -x -q -m gpu -k "cuda_graph" > gpurun_out/pytest_graph2.log 2>&1; tail -15 gpurun_out/pytest_graph2.log
