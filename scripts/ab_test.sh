cp concrete-numpy_b200/libfhe_b200.so /tmp/base.so
echo "== BASE"; timeout 100 python scripts/quick_bench.py p5 p4 2>&1 | tail -2
for v in b c; do cp scripts/ubench/libfhe_var_$v.so concrete-numpy_b200/libfhe_b200.so; echo "== VAR $v"; timeout 100 python scripts/quick_bench.py p5 2>&1 | tail -1; done
cp /tmp/base.so concrete-numpy_b200/libfhe_b200.so
