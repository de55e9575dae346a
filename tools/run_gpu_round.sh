set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu15.log 2>&1; tail -3 gpurun_out/pytest_gpu15.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1i.log 2>&1; tail -1 gpurun_out/bench_r1i.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1i.csv python tools/profile_step.py --batch 64 > gpurun_out/ncu_step_r1i.log 2>&1; tail -2 gpurun_out/ncu_step_r1i.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 2 -c 1 -f -o gpurun_out/prof_res_r1i python tools/conv_bench.py --only "G.res" --reps 1 > gpurun_out/ncu_res_r1i.log 2>&1; tail -2 gpurun_out/ncu_res_r1i.log
