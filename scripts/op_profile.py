"""Development aid: per-op-type device+host time of one circuit run (synchronising after every op)."""
import os, sys, time, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from concrete_numpy_b200 import api, executor as X
from golden_util import load_golden, to_value
name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_kvdb128_insert"
doc = load_golden(name)
circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], api.Configuration(jit=True, global_p_error=doc["global_p_error"]))
args = [to_value(a) for a in doc["cases"][0]["args"]]
circuit.keygen()
enc = circuit.encrypt(*args)
ev = circuit.client.evaluation_keys
circuit.server.run(enc, ev)
acc = collections.defaultdict(lambda: [0, 0.0, 0])
orig_init = X.Executor.run
def timed_run(self, evk, a):
    for k, h in list(self.handlers.items()):
        def mk(k=k, h=h):
            def w(op, v):
                torch.cuda.synchronize(); t0 = time.perf_counter()
                r = h(op, v)
                torch.cuda.synchronize(); dt = time.perf_counter() - t0
                acc[k][0] += 1; acc[k][1] += dt; acc[k][2] += (op.result_type.size if hasattr(op.result_type, "size") else 0)
                return r
            return w
        self.handlers[k] = mk()
    return orig_init(self, evk, a)
X.Executor.run = timed_run
torch.cuda.synchronize(); t0 = time.perf_counter()
circuit.server.run(enc, ev)
torch.cuda.synchronize(); total = time.perf_counter() - t0
print(name, "total %.1f ms" % (total * 1e3))
for k, (n, t, sz) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"  {k:28s} calls {n:5d}  {t*1e3:9.2f} ms  ({t/n*1e6:8.1f} us/call, {sz} result elems)")
