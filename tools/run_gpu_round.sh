set -x
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_r1x.log 2>&1; tail -1 gpurun_out/smoke_r1x.log | cut -c1-160
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu32.log 2>&1; tail -2 gpurun_out/pytest_gpu32.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r1x.log 2>&1; tail -1 gpurun_out/bench_r1x.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1x.csv python tools/profile_step.py --batch 64 > gpurun_out/ncu_step_r1x.log 2>&1; tail -1 gpurun_out/ncu_step_r1x.log
