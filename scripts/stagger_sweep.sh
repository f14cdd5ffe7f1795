for s in 0 9500 12600 19000; do echo "== stagger $s"; FHE_PBS_STAGGER=$s timeout 100 python scripts/quick_bench.py p8 2>&1 | tail -1; done
FHE_PBS_STAGGER=19000 timeout 100 python scripts/phase_prof.py 8 2>&1 | tail -10
