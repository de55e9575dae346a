set -x
timeout 900 python -m pytest tests/test_batchnorm_gpu.py -x -q -m gpu > gpurun_out/pytest_bn.log 2>&1; tail -30 gpurun_out/pytest_bn.log
