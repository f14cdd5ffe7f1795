"""torchrun target: one circuit sharded by ciphertext over all ranks (NCCL all-gather after every table
lookup), checked against the reference's clear evaluation and timed against the unsharded run.

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/dist_check.py [names...]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from concrete_numpy_b200 import api, compiler as cc
from golden_util import load_golden, same, to_value

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{local}")
torch.cuda.set_device(dev)
os.environ["CNP_B200_DEVICE"] = str(local)
os.environ.setdefault("CNP_B200_KEY_SEED", "20261019")      # every rank derives the same keys
dist.init_process_group("nccl", device_id=dev)
names = sys.argv[1:] or ["cfg3_matmul6_relu", "cfg5_kvdb128_query", "cfg1_lut4_1024"]
res = {}
for name in names:
    doc = load_golden(name)
    circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"],
                          api.Configuration(jit=True, global_p_error=doc["global_p_error"]))
    case = doc["cases"][0]
    args = [to_value(a) for a in case["args"]]
    circuit.keygen()
    enc = circuit.encrypt(*args)
    ev = circuit.client.evaluation_keys

    def timed():
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = circuit.server.run(enc, ev)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t.item())

    cc.set_shard_group(None)
    timed()
    out1, ms1 = timed()
    cc.set_shard_group(dist.group.WORLD)
    timed()
    outN, msN = timed()
    ok1 = same(circuit.decrypt(out1), to_value(case["expected"]))
    okN = same(circuit.decrypt(outN), to_value(case["expected"]))
    okt = torch.tensor([int(ok1 and okN)], device=dev)
    dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    res[name] = {"world": world, "correct_all_ranks": bool(okt.item()), "unsharded_ms": round(ms1, 2),
                 "sharded_ms": round(msN, 2), "speedup": round(ms1 / msN, 3)}
    if rank == 0:
        print(name, json.dumps(res[name]), flush=True)
    cc.set_shard_group(None)
    circuit.cleanup()
if rank == 0:
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open(f"gpurun_out/dist_check_{world}gpu.json", "w"), indent=1)
dist.destroy_process_group()
