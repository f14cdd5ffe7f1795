cp concrete-numpy_b200/libfhe_b200.so /tmp/base.so
echo "== BASE"; timeout 200 python scripts/quick_bench.py p8 p6 2>&1 | tail -2
for v in a; do cp scripts/ubench/libfhe_var_$v.so concrete-numpy_b200/libfhe_b200.so; echo "== VAR $v"; timeout 200 python scripts/quick_bench.py p8 p6 2>&1 | tail -2; done
cp /tmp/base.so concrete-numpy_b200/libfhe_b200.so
echo "== BASE again"; timeout 200 python scripts/quick_bench.py p8 2>&1 | tail -1
