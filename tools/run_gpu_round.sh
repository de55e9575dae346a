set -x
timeout 900 python -m pytest tests/test_conv_umma_gpu.py tests/test_ops_gpu.py -x -q -m gpu > gpurun_out/pytest_gpu14.log 2>&1; tail -3 gpurun_out/pytest_gpu14.log
timeout 600 python tools/conv_bench.py > gpurun_out/cb_tail.log 2>&1; grep -E "^G\.|^D1\.[23]|total|top" gpurun_out/cb_tail.log
SRGAN_DBG_WGRAD_PW=16 timeout 600 python tools/conv_bench.py --only "G." > gpurun_out/cb_pw16.log 2>&1; grep -E "wgrad" gpurun_out/cb_pw16.log
SRGAN_DBG_CONV_TAIL=0 timeout 600 python tools/conv_bench.py --only "G.res" > gpurun_out/cb_notail.log 2>&1; grep -E "G.res" gpurun_out/cb_notail.log
