set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu16.log 2>&1; tail -3 gpurun_out/pytest_gpu16.log
timeout 600 python tools/step_trace.py --batch 64 > gpurun_out/step_trace_r1j.log 2>&1; head -60 gpurun_out/step_trace_r1j.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1j.log 2>&1; tail -1 gpurun_out/bench_r1j.log | cut -c1-300
