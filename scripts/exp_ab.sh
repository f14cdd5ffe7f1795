#!/bin/bash
# Development aid: A/B of a k_pbs build knob on one B200 (usage: scripts/exp_ab.sh ENVVAR "sets" [pytest -k expr])
# Runs the bootstrap parity tests with the default setting, then quick_bench + phase_prof with ENVVAR=0 and =1.
VAR=${1:-FHE_PBS_ACCTM}; SETS=${2:-"p8"}; KEXPR=${3:-"pbs or polymul"}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "$KEXPR" > gpurun_out/ab_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/ab_pytest.log
tail -3 gpurun_out/ab_pytest.log
for v in 0 1; do
  echo "== $VAR=$v quick_bench $SETS"
  env $VAR=$v python scripts/quick_bench.py $SETS 2>&1 | grep -v "^{'fp64" | tee gpurun_out/ab_quick_$v.log
  cp gpurun_out/quick_bench.json gpurun_out/ab_quick_$v.json
  for s in $SETS; do
    echo "== $VAR=$v phase_prof ${s#p}"
    env $VAR=$v python scripts/phase_prof.py ${s#p} 2>&1 | tee gpurun_out/ab_phase_${s}_$v.log
  done
done
