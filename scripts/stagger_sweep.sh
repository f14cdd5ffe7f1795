for s in 0 1400 2800 5600; do echo "== stagger $s"; FHE_PBS_STAGGER=$s python scripts/quick_bench.py p8 2>&1 | tail -1; done
FHE_PBS_STAGGER=2800 python scripts/phase_prof.py 8 2>&1 | tail -10
