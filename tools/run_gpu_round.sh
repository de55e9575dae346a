set -x
timeout 300 python -c "import __graft_entry__ as g; g.build(); g.smoke()" > gpurun_out/smoke_r1t.log 2>&1; tail -2 gpurun_out/smoke_r1t.log
timeout 600 python bench.py --steps 8 --warmup 3 > gpurun_out/bench_r1t.log 2>&1; tail -1 gpurun_out/bench_r1t.log | cut -c1-200
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r1t.log 2>&1; tail -1 gpurun_out/bench_ref_r1t.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r1t.csv python tools/profile_step.py --batch 64 > gpurun_out/ncu_step_r1t.log 2>&1; tail -1 gpurun_out/ncu_step_r1t.log
timeout 300 python tools/conv_bench.py > gpurun_out/cb_r1t.log 2>&1; grep -E "total" gpurun_out/cb_r1t.log
