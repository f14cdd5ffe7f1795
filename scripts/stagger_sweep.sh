for s in 0 1000 4000 8000; do echo "== stagger $s"; FHE_PBS_STAGGER=$s python scripts/phase_prof.py 8 2>&1 | tail -11; done
for s in 0 4000; do echo "== stagger $s quick"; FHE_PBS_STAGGER=$s python scripts/quick_bench.py p8 p4 2>&1 | tail -2; done
