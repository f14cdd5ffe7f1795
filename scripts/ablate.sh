for d in 0 1 2 4 8 15; do echo "== DBG $d"; FHE_PBS_DBG=$d python scripts/phase_prof.py 8 2>&1 | tail -10; done
