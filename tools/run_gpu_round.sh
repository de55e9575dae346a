set -x
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -k "instance_norm" > gpurun_out/pytest_norm.log 2>&1; tail -5 gpurun_out/pytest_norm.log
timeout 300 python tools/conv_bench.py --only IN > gpurun_out/norm_resident.log 2>&1; cat gpurun_out/norm_resident.log
