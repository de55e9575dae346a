set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu19.log 2>&1; tail -3 gpurun_out/pytest_gpu19.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_r1o.log 2>&1; tail -1 gpurun_out/bench_r1o.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload srgan_nb05 --batch 32 > gpurun_out/bench_r1o_nb05.log 2>&1; tail -1 gpurun_out/bench_r1o_nb05.log | cut -c1-200
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --workload nb02_solo > gpurun_out/bench_r1o_nb02.log 2>&1; tail -1 gpurun_out/bench_r1o_nb02.log | cut -c1-200
