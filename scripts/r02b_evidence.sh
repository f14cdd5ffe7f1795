set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02b_gputest.log 2>&1; tail -3 gpurun_out/r02b_gputest.log
python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err; tail -c 300 gpurun_out/r02b_bench.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02b_ncu_bench.log 2>&1; tail -2 gpurun_out/r02b_ncu_bench.log | cut -c1-200
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_pbs_group -c 1 -o gpurun_out/r02b_pbs_group_p4 -f python scripts/prof_pbs.py p4 1184 > gpurun_out/r02b_ncu_group.log 2>&1; tail -2 gpurun_out/r02b_ncu_group.log
python scripts/quick_bench.py p8 p6 p5 p4 > gpurun_out/r02b_quick_bench.log 2>&1; tail -4 gpurun_out/r02b_quick_bench.log
python scripts/phase_prof.py 4 > gpurun_out/r02b_phase_p4.log 2>&1
python scripts/group_ablation.py > gpurun_out/r02b_group_ablation.log 2>&1; cat gpurun_out/r02b_group_ablation.log
python scripts/config_latency.py > gpurun_out/r02b_config_latency.log 2>&1; cp gpurun_out/config_latency.json gpurun_out/r02b_config_latency.json
python scripts/variant_latency.py p5 p4 > gpurun_out/r02b_variant_latency.log 2>&1; tail -12 gpurun_out/r02b_variant_latency.log
