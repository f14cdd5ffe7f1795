nvidia-smi --query-gpu=clocks.sm,power.draw,temperature.gpu --format=csv,noheader
echo "== NEW"; timeout 200 python scripts/quick_bench.py p4 p8 2>&1 | tail -2
cp concrete-numpy_b200/libfhe_b200.so /tmp/new.so; cp scripts/ubench/libfhe_b200_old.so concrete-numpy_b200/libfhe_b200.so
echo "== OLD"; timeout 200 python scripts/quick_bench.py p4 p8 2>&1 | tail -2
cp /tmp/new.so concrete-numpy_b200/libfhe_b200.so
echo "== NEW again"; timeout 200 python scripts/quick_bench.py p4 2>&1 | tail -1
