for d in 0 6 14; do echo "== DBG $d"; FHE_PBS_DBG=$d timeout 100 python scripts/phase_prof.py 8 2>&1 | tail -11; done
