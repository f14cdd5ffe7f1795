for s in 1500 3000 6000; do echo "== SMALL_L 1 stagger $s"; FHE_PBS_STAGGER=$s FHE_PBS_SMALL_L=1 python scripts/quick_bench.py p8 2>&1 | tail -1; done
FHE_PBS_STAGGER=3000 FHE_PBS_SMALL_L=1 python scripts/quick_bench.py p6 2>&1 | tail -1
FHE_PBS_STAGGER=3000 FHE_PBS_SMALL_L=1 python scripts/phase_prof.py 8 2>&1 | tail -11
