"""Development aid: device time per operation type of one `Server.run` of a golden circuit (CUDA events around every
executor handler; the gaps the host leaves between kernels count towards the operation that follows them)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from concrete_numpy_b200 import api
from golden_util import load_golden, to_value

name = sys.argv[1] if len(sys.argv) > 1 else "cfg5_kvdb128_insert"
doc = load_golden(name)
circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], api.Configuration(jit=True, global_p_error=doc["global_p_error"]))
args = [to_value(a) for a in doc["cases"][0]["args"]]
circuit.keygen()
enc = circuit.encrypt(*args)
ev = circuit.client.evaluation_keys
circuit.server.run(enc, ev)
circuit.server.run(enc, ev)
ex = list(circuit.server._server_lambda._executors.values())[0]
events = []
for base, h in list(ex.handlers.items()):
    def wrapped(op, v, _h=h, _b=base):
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
        out = _h(op, v)
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        n = op.result_type.size if hasattr(op.result_type, "size") else 0
        events.append((_b, n, e0, e1))
        return out
    ex.handlers[base] = wrapped
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); t0.record()
circuit.server.run(enc, ev)
t1.record(); torch.cuda.synchronize()
tot = t0.elapsed_time(t1)
agg = collections.defaultdict(lambda: [0, 0.0])
inside = 0.0
for b, n, e0, e1 in events:
    ms = e0.elapsed_time(e1); inside += ms
    a = agg[(b, n)]; a[0] += 1; a[1] += ms
gaps = sum(events[i][3].elapsed_time(events[i + 1][2]) for i in range(len(events) - 1))
print(f"{name}: run {tot:.1f} ms, {len(events)} operations, inside handlers {inside:.1f} ms, between handlers {gaps:.1f} ms")
for (b, n), (c, ms) in sorted(agg.items(), key=lambda t: -t[1][1])[:14]:
    print(f"   {b:28s} size {n:6d}  x{c:5d}  {ms:9.2f} ms  ({ms / c * 1e3:8.1f} us each)")
