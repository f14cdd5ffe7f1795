set -x
python -m pytest tests -x -q -m gpu > gpurun_out/r02_gputest_final.log 2>&1; tail -3 gpurun_out/r02_gputest_final.log
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 300 gpurun_out/r02_bench_final.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-configs --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1; tail -2 gpurun_out/r02_ncu_bench.log | cut -c1-200
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_pbs_pipe -c 1 --csv --log-file gpurun_out/r02_pbs_p8_dram_full_launch.csv python scripts/prof_pbs.py p8 4096 > gpurun_out/r02_ncu_dram.log 2>&1; tail -3 gpurun_out/r02_pbs_p8_dram_full_launch.csv | cut -c1-300
timeout 800 ncu --set full --import-source on --clock-control none -k regex:k_pbs_pipe -c 1 -o gpurun_out/r02_pbs_pipe_final -f python scripts/prof_pbs.py p8 30 > gpurun_out/r02_ncu_full.log 2>&1; tail -2 gpurun_out/r02_ncu_full.log
python scripts/quick_bench.py p8 p6 p5 p4 > gpurun_out/r02_quick_bench_final.log 2>&1; tail -4 gpurun_out/r02_quick_bench_final.log
python scripts/phase_prof.py 8 6 > gpurun_out/r02_phase_final.log 2>&1
python scripts/op_profile.py > gpurun_out/r02_op_profile_final.log 2>&1; head -3 gpurun_out/r02_op_profile_final.log
scripts/ablate_pipe.sh > gpurun_out/r02_pipe_ablation_final.log 2>&1
