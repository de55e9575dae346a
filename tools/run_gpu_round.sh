set -x
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu21.log 2>&1; tail -2 gpurun_out/pytest_gpu21.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r1q.log 2>&1; tail -1 gpurun_out/bench_r1q.log | cut -c1-150
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1q.csv python tools/profile_step.py --batch 64 > gpurun_out/ncu_step_r1q.log 2>&1; tail -1 gpurun_out/ncu_step_r1q.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_umma_kernel -s 2 -c 1 -f -o gpurun_out/prof_res_r1q python tools/conv_bench.py --only "G.res" --reps 1 > gpurun_out/ncu_res_r1q.log 2>&1; tail -1 gpurun_out/ncu_res_r1q.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_umma_kernel -s 2 -c 1 -f -o gpurun_out/prof_wgrad_r1q python tools/conv_bench.py --only "G.res" --reps 1 > gpurun_out/ncu_wgrad_r1q.log 2>&1; tail -1 gpurun_out/ncu_wgrad_r1q.log
timeout 300 ncu --set full --clock-control none -k regex:inorm_apply_kernel -s 2 -c 1 -f -o gpurun_out/prof_inorm_r1q python tools/conv_bench.py --only "IN 256@32" --reps 1 > gpurun_out/ncu_inorm_r1q.log 2>&1; tail -1 gpurun_out/ncu_inorm_r1q.log
