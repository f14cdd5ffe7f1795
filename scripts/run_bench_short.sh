timeout 200 python -m pytest tests/test_golden_gpu.py -m gpu -q -k "cfg1 or cfg3 or cfg2" 2>&1 | tail -3
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu-baseline 2> gpurun_out/bench_pipe.err | tee gpurun_out/bench_pipe.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['roofline']['share_of_step'], d['e2e']['value'], d['config']['decrypted_output_correct'])"
tail -3 gpurun_out/bench_pipe.err
