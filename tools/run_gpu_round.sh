set -x
timeout 600 python -m pytest tests/test_step_gpu.py -x -q -m gpu -k "checkpoint or cuda_graph" > gpurun_out/pytest_ckpt.log 2>&1; tail -25 gpurun_out/pytest_ckpt.log
