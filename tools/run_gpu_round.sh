set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu18.log 2>&1; tail -3 gpurun_out/pytest_gpu18.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1l.log 2>&1; tail -1 gpurun_out/bench_r1l.log
timeout 600 python bench.py --steps 5 --warmup 3 --graph off --no-cpu > gpurun_out/bench_r1l_eager.log 2>&1; tail -1 gpurun_out/bench_r1l_eager.log | cut -c1-200
for b in 8 16 32; do timeout 600 python bench.py --steps 5 --warmup 3 --batch $b --no-cpu > gpurun_out/bench_r1l_b$b.log 2>&1; tail -1 gpurun_out/bench_r1l_b$b.log | cut -c1-200; timeout 600 python bench.py --steps 5 --warmup 3 --batch $b --no-cpu --graph off > gpurun_out/bench_r1l_b${b}_eager.log 2>&1; tail -1 gpurun_out/bench_r1l_b${b}_eager.log | cut -c1-200; done
