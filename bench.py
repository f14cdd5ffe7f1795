#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200 encrypted-execution backend.

    python bench.py --gpus N --steps K --warmup W            # this backend
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, host cores)

Metric (BASELINE.json): device-timed PBS/s.  Workload: BASELINE config 2 -- the 8-bit univariate
`(x + 3) ** 2 // 5` on a 64x64 encrypted tensor (two chained table lookups = 8192 keyswitch+bootstrap per
step) with the published 8-bit parameter set (n=996, k=1, N=32768, pbs level 2 / base 2^15, ks level 7 /
base 2^3).  A step = one `Server.run` of the circuit.

  value : PBS/s with the input ciphertexts already resident in HBM (CUDA events, max over ranks)
  e2e   : PBS/s through the public API (`Server.run`) with HOST ciphertext buffers: pinned host ->
          device copy of the arguments and device -> host copy of the result inside the timed region
  roofline : dominant kernel = k_pbs, fp64 (DFMA) bound; achieved = algorithmic flops per launch
          (BASELINE.md section 4 formula) / mean k_pbs launch time (CUDA events on the launch stream);
          peak = DFMA throughput measured live on this GPU (MEASURED_PEAKS.json has no fp64 figure);
          traffic = DRAM bytes of one launch from the ncu capture recorded in profiles/ (file + commit named)
  cpu_baseline : the oracle (C port of the reference algorithm, scalar radix-2 FFT) on the box's host cores,
          on a bounded sample of the same workload
  configs (N = 1) : the other BASELINE configs (cfg1, cfg3, cfg4, cfg5 query / replace / insert): device-timed
          `Server.run` latency, PBS/s, decrypted-output check, the CPU port's PBS/s at THAT config's parameter
          set on a bounded sample, and for cfg3 / cfg4 the HBM fraction of the matmul / conv2d kernel
  strong_scaling (N > 1) : ONE circuit sharded by ciphertext over the N ranks (executor.py) for cfg2,
          cfg1, cfg3, cfg4 and cfg5 query / insert: latency (max over ranks), the single-GPU latency measured in the
          same run, efficiency = T1 / (N * TN), and the bytes / device time of the collectives
Multi-GPU headline (`value`): ciphertexts are independent -> each rank runs the circuit on its own 64x64 batch
with replicated keys (deterministic keygen from one seed, no key broadcast, no data-path collective); weak
scaling, value = all ranks' PBS / max-over-ranks time.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "cfg2_univariate8_64x64"
CPU_SAMPLE_PBS = 16
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_pbs_p8_traffic.json")
OTHER_CONFIGS = ["cfg1_lut4_1024", "cfg3_matmul6_relu", "cfg4_conv2d_relu", "cfg5_kvdb128_query",
                 "cfg5_kvdb128_replace", "cfg5_kvdb128_insert"]
SHARDED_CONFIGS = ["cfg1_lut4_1024", "cfg3_matmul6_relu", "cfg4_conv2d_relu", "cfg5_kvdb128_query",
                   "cfg5_kvdb128_insert"]


def golden(name=WORKLOAD):
    from golden_util import load_golden

    return load_golden(name)


def flops_per_pbs(p):
    M = p.N // 2
    return p.n * ((p.k + 1) * (p.pbs_level + 1) * 5 * M * np.log2(M) + 8 * (p.k + 1) ** 2 * p.pbs_level * M)


def workload_config(params):
    """Static description of the headline workload, printed identically by both arms."""
    return {"workload": WORKLOAD, "circuit": "(x + 3) ** 2 // 5 on 64x64 eint<8>", "pbs_per_step_per_gpu": 8192,
            "params": params.to_dict(), "sharding": "by ciphertext, replicated BSK/KSK, no collective",
            "l2_flush": "inputs (1.07 GB of ciphertexts per step) exceed the 126 MB L2"}


def params_brief(p):
    return {"p": p.p, "n": p.n, "k": p.k, "N": p.N, "pbs_level": p.pbs_level, "pbs_base_log": p.pbs_base_log,
            "ks_level": p.ks_level, "ks_base_log": p.ks_base_log}


class ClockSampler:
    """nvidia-smi sampler running during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for line in self.lines:
            f = [t.strip() for t in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ CPU arm --------------------------------

def host_threads():
    """All host threads for the OpenMP port.  torchrun exports OMP_NUM_THREADS=1 to its workers, which would run
    the CPU arm single-threaded: set the count explicitly, in the environment (read when libgomp loads) and on
    the runtime if it is already loaded."""
    cores = os.cpu_count() or 1
    try:
        cores = len(os.sched_getaffinity(0)) or cores
    except (AttributeError, OSError):
        pass
    os.environ["OMP_NUM_THREADS"] = str(cores)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    try:
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    return cores


def cpu_sample(params, n_pbs, native=True):
    """Oracle timing on the host cores: keyswitch + bootstrap of `n_pbs` ciphertexts with synthetic
    (bounded random) key material -- the run time of the port does not depend on key values."""
    from oracle.oracle import Oracle

    try:
        o = Oracle(native=native)
    except Exception:                                                     # noqa: BLE001
        o = Oracle(native=False)
    rng = np.random.default_rng(0)
    p = params
    chunk = rng.integers(0, 1 << 64, 1 << 16, dtype=np.uint64)
    ksk = np.resize(chunk, (p.big_dim, p.ks_level, p.n + 1))
    bsk_f = np.resize(rng.standard_normal(1 << 16) * 2.0 ** 40, (p.n, p.pbs_level, p.k + 1, p.k + 1, p.N // 2, 2))
    ct = rng.integers(0, 1 << 64, (n_pbs, p.big_dim + 1), dtype=np.uint64)
    lut = np.zeros((1, (p.k + 1) * p.N), dtype=np.uint64)
    lut[0, p.k * p.N:] = o.build_lut_poly(np.arange(1 << p.p) % (1 << p.p), p.p, p.p, p.N)
    t0 = time.perf_counter()
    small = o.keyswitch(ct, ksk, p.ks_level, p.ks_base_log)
    o.pbs(small, bsk_f, lut, None, p.pbs_base_log)
    dt = time.perf_counter() - t0
    return n_pbs / dt, dt


CPU_KIND_NOTE = "oracle/tfhe_ref.c: C port of the reference algorithm, scalar radix-2 fp64 FFT, OpenMP over ciphertexts"


def run_reference(args):
    """CPU arm: the oracle port of the reference algorithm on all host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = host_threads()
    from concrete_numpy_b200.params import REFERENCE_SETS

    params = REFERENCE_SETS[8]
    # a step = a bounded sample of the workload: two ciphertexts per thread (>= 16), about 2-4 s on the port,
    # so that steps + warm-up end within a few minutes
    n = max(CPU_SAMPLE_PBS, 2 * cores)
    for _ in range(min(args.warmup, 1)):
        cpu_sample(params, n)
    times, count = [], 0
    for _ in range(args.steps):
        _, dt = cpu_sample(params, n)
        times.append(dt)
        count += n
    total = sum(times)
    value = count / total
    line = {
        "impl": "reference", "metric": "PBS/s (device-timed)", "value": value, "unit": "PBS/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(params),
        "cpu_baseline": {"value": value, "unit": "PBS/s", "cores": cores, "kind": "port", "fft": "scalar radix-2",
                         "sample": f"{n} keyswitch+bootstrap of the same parameter set per step, synthetic BSK/KSK; {CPU_KIND_NOTE}"},
        "e2e": {"value": value, "unit": "PBS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ GPU arm --------------------------------

def load_traffic():
    try:
        with open(TRAFFIC_FILE, "r", encoding="utf-8") as f:
            return json.load(f)
    except (OSError, ValueError):
        return None


class LinearTimer:
    """CUDA events around the matmul / gather-MAC (conv2d, sum) launches of a run, with their algorithmic bytes
    (matmul (M*K + M*N)*L*8 + K*N*8, conv (in + out)*L*8)."""

    def __init__(self, E, torch):
        self.E, self.torch, self.records = E, torch, []
        self._mm, self._gm = E.matmul_ct_clear, E.lin_gather_mac

    def __enter__(self):
        torch, E = self.torch, self.E

        def timed(name, fn, nbytes, *a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = fn(*a, **k)
            e1.record()
            self.records.append((name, e0, e1, nbytes))
            return out

        def mm(x, W, Mr, K, Nc, L):
            return timed("k_matmul_taps (tap lists + dense fallback launch)", self._mm, (Mr * K + Mr * Nc) * L * 8 + K * Nc * 8, x, W, Mr, K, Nc, L)

        def gm(x, idx, w, L, bias=None):
            return timed("k_gather_mac", self._gm, (x.shape[0] + idx.shape[0]) * L * 8, x, idx, w, L, bias)

        E.matmul_ct_clear, E.lin_gather_mac = mm, gm
        return self

    def __exit__(self, *exc):
        self.E.matmul_ct_clear, self.E.lin_gather_mac = self._mm, self._gm

    def summary(self, hbm_gbs):
        self.torch.cuda.synchronize()
        best = None
        for name, e0, e1, nbytes in self.records:
            ms = e0.elapsed_time(e1)
            if best is None or nbytes > best["algorithmic_bytes"]:
                gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
                best = {"kernel": name, "ms": ms, "algorithmic_bytes": int(nbytes), "achieved_gbs": gbs,
                        "peak_gbs": hbm_gbs, "frac": gbs / hbm_gbs if hbm_gbs else None}
        return best


def run_ours(args):
    import torch
    import torch.distributed as dist

    from concrete_numpy_b200 import _lib, api, compiler as cc, engine as E
    from concrete_numpy_b200.executor import Ct
    from concrete_numpy_b200.params import REFERENCE_SETS
    from golden_util import same, to_value

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner there at every debug level the box may
        # preset (VERSION / WARN); send its diagnostics to stderr's neighbour instead unless the user asked for a file
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION", "WARN"):
            os.environ["NCCL_DEBUG"] = "WARN"
            os.environ.setdefault("NCCL_DEBUG_FILE", "/tmp/nccl_bench_%h_%p.log")
        dist.init_process_group("nccl", device_id=dev)

    doc = golden()
    params = REFERENCE_SETS[8]
    circuit = api.Circuit(doc["mlir"], doc["input_signs"], doc["output_signs"], api.Configuration(jit=True),
                          crypto_parameters=params)
    os.environ["CNP_B200_DEVICE"] = str(local)
    os.environ.setdefault("CNP_B200_KEY_SEED", "42")       # every rank derives the same keys (no key broadcast)
    keyset = cc.ClientSupport.key_set(circuit.server.client_specs.client_parameters, seed=42, device=dev)
    circuit.client._keyset = keyset
    ev = keyset.get_evaluation_keys()

    rng = np.random.default_rng(rank)
    x = rng.integers(0, 13, (64, 64))
    enc = circuit.client.encrypt(x)                     # PublicArguments with a device Ct
    ct_dev = enc.values[0]
    host_in = torch.empty(ct_dev.data.shape, dtype=torch.int64, pin_memory=True)
    host_in.copy_(ct_dev.data)
    host_out = torch.empty(ct_dev.data.shape, dtype=torch.int64, pin_memory=True)
    n_pbs_step = circuit.server._feedback.n_pbs
    fp64_peak = E.measure_fp64_peak(dev)

    # instrument the bootstrap launches with CUDA events (torch's current stream = launch stream)
    pbs_events = []
    raw_pbs_launch = E.pbs_launch

    def timed_pbs_launch(ev_, out_, small_, count, luts_, idx_, *rest):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        raw_pbs_launch(ev_, out_, small_, count, luts_, idx_, *rest)
        e1.record()
        pbs_events.append((e0, e1, count))

    E.pbs_launch = timed_pbs_launch

    def step_resident():
        return circuit.server.run(enc, ev)

    def step_e2e():
        d = host_in.to(dev, non_blocking=True)
        res = circuit.server.run(cc.PublicArguments([Ct(d, ct_dev.shape)]), ev)
        host_out.copy_(res.values[0].data, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # correctness of what is being timed (decrypted result == reference clear evaluation)
    cc.set_shard_group(None)
    res = step_resident()
    got = circuit.client.decrypt(res)
    correct = bool(np.array_equal(got, (x + 3) ** 2 // 5))
    for _ in range(max(args.warmup - 1, 0)):
        step_resident()
    pbs_events.clear()
    launches0 = _lib.launches_total()
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step_resident()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = _lib.launches_total() - launches0
    pbs_ms = sum(a.elapsed_time(b) for a, b, _ in pbs_events)
    pbs_cts = sum(n for _, _, n in pbs_events)
    n_pbs_launch = len(pbs_events)

    # end to end through Server.run with host buffers
    step_e2e()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        step_e2e()
    g1.record()
    barrier()
    e2e_ms = g0.elapsed_time(g1)
    E.pbs_launch = raw_pbs_launch

    if world > 1:
        t = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        ok = torch.tensor([1 if correct else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        correct = bool(ok.item())

    hbm_gbs = None
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json"), "r", encoding="utf-8") as f:
            hbm_gbs = float(json.load(f)["hbm_gbs"])
    except (OSError, ValueError, KeyError):
        hbm_gbs = 6650.0            # B200_PROFILING.md fallback

    def time_run(circ, enc_, ev_, reps=3, sync_ranks=False):
        """median device time of `Server.run` (CUDA events; max over ranks when the run is sharded)"""
        times = []
        for _ in range(reps):
            torch.cuda.synchronize()
            if sync_ranks:
                dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out_ = circ.server.run(enc_, ev_)
            b.record()
            torch.cuda.synchronize()
            t_ms = a.elapsed_time(b)
            if sync_ranks:
                tt = torch.tensor([t_ms], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                t_ms = float(tt.item())
            times.append(t_ms)
        return float(np.median(times)), out_

    def open_config(name):
        d = golden(name)
        circ = api.Circuit(d["mlir"], d["input_signs"], d["output_signs"],
                           api.Configuration(jit=True, global_p_error=d["global_p_error"]))
        case = d["cases"][0]
        circ.keygen()
        enc_ = circ.encrypt(*[to_value(a) for a in case["args"]])
        return circ, enc_, circ.client.evaluation_keys, to_value(case["expected"])

    configs, strong = None, None
    if world == 1 and not args.no_configs:
        # ---- the other BASELINE configs on this GPU, each next to the CPU port at its own parameter set
        configs = {}
        cores = host_threads() if not args.no_cpu_baseline else None
        for name in OTHER_CONFIGS:
            circ, enc_, ev_, expected = open_config(name)
            out_ = circ.server.run(enc_, ev_)                       # warm-up + correctness
            ok_ = bool(same(circ.decrypt(out_), expected))
            lin = None
            if name.startswith(("cfg3", "cfg4")):
                with LinearTimer(E, torch) as lt:
                    circ.server.run(enc_, ev_)
                    lin = lt.summary(hbm_gbs)
            t_ms, _ = time_run(circ, enc_, ev_)
            p_ = circ.server.client_specs.client_parameters.crypto
            n_pbs_ = int(circ.server._feedback.n_pbs)
            entry = {"latency_ms": t_ms, "n_pbs": n_pbs_, "pbs_per_s": n_pbs_ / (t_ms * 1e-3),
                     "decrypted_output_correct": ok_, "params": params_brief(p_)}
            if fp64_peak:
                # whole-circuit fp64 roofline fraction: the bootstraps' algorithmic flops over the circuit's
                # device time (keyswitches, linear kernels and partial waves included) over the measured DFMA peak
                tf_ = n_pbs_ * flops_per_pbs(p_) / (t_ms * 1e-3) / 1e12
                entry["roofline"] = {"bound": "fp64", "achieved": tf_, "peak": fp64_peak, "unit": "TFLOP/s",
                                     "frac": tf_ / fp64_peak, "flops_per_pbs": flops_per_pbs(p_)}
            if lin is not None and name.startswith(("cfg3", "cfg4")):
                entry["linear_kernel"] = lin
            if cores:
                n_s = max(CPU_SAMPLE_PBS, cores)
                v, dt = cpu_sample(p_, n_s)
                entry["cpu_baseline"] = {"value": v, "unit": "PBS/s", "cores": cores, "kind": "port", "fft": "scalar radix-2",
                                         "sample": f"{n_s} keyswitch+bootstrap at this parameter set in {dt:.2f} s",
                                         "latency_estimate_ms": 1e3 * n_pbs_ / v}
            configs[name] = entry
            circ.cleanup()
            del circ, enc_, ev_, out_
            torch.cuda.empty_cache()
    if world > 1 and not args.no_configs:
        # ---- one circuit sharded over the ranks (strong scaling): same ciphertexts on every rank
        strong = {}

        def sharded(circ, enc_, ev_, expected, t1_ms=None, reps=3):
            for v_ in enc_.values:                                  # every rank must hold rank 0's ciphertexts
                if isinstance(v_, Ct):
                    dist.broadcast(v_.data, src=0)
            if t1_ms is None:
                cc.set_shard_group(None)
                circ.server.run(enc_, ev_)
                t1_ms, _ = time_run(circ, enc_, ev_, reps=reps, sync_ranks=True)
            cc.set_shard_group(dist.group.WORLD)
            ex = circ.server._server_lambda.executor(dev)
            out_ = circ.server.run(enc_, ev_)
            ok_t = torch.tensor([1 if same(circ.decrypt(out_), expected) else 0], device=dev)
            dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
            tn_ms, _ = time_run(circ, enc_, ev_, reps=reps, sync_ranks=True)
            ex.time_comm = True
            circ.server.run(enc_, ev_)
            st = dict(ex.last_stats)
            ex.time_comm = False
            cc.set_shard_group(None)
            comm = torch.tensor([st.get("comm_ms", 0.0)], device=dev, dtype=torch.float64)
            dist.all_reduce(comm, op=dist.ReduceOp.MAX)
            return {"latency_ms": tn_ms, "single_gpu_latency_ms": t1_ms, "speedup": t1_ms / tn_ms,
                    "efficiency": t1_ms / (world * tn_ms), "correct_all_ranks": bool(ok_t.item()),
                    "n_pbs": int(st.get("n_pbs", 0)), "allgather_calls": st.get("allgather_calls", 0),
                    "allgather_bytes": st.get("allgather_bytes", 0), "allreduce_calls": st.get("allreduce_calls", 0),
                    "allreduce_bytes": st.get("allreduce_bytes", 0), "collectives_ms": float(comm.item())}

        # cfg2: rank 0's input is broadcast, so every rank checks against rank 0's clear result
        x0 = np.random.default_rng(0).integers(0, 13, (64, 64))
        strong[WORKLOAD] = sharded(circuit, enc, ev, (x0 + 3) ** 2 // 5, t1_ms=ms / args.steps, reps=2)
        for name in SHARDED_CONFIGS:
            circ, enc_, ev_, expected = open_config(name)
            strong[name] = sharded(circ, enc_, ev_, expected)
            circ.cleanup()
            del circ, enc_, ev_
            torch.cuda.empty_cache()

    if rank == 0:
        total_pbs = n_pbs_step * args.steps * world
        value = total_pbs / (ms * 1e-3)
        e2e_value = total_pbs / (e2e_ms * 1e-3)
        fl = flops_per_pbs(params)
        achieved = pbs_cts * fl / (pbs_ms * 1e-3) / 1e12 if pbs_ms > 0 else 0.0
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = host_threads()
            n = max(CPU_SAMPLE_PBS, cores)
            v, dt = cpu_sample(params, n)
            cpu = {"value": v, "unit": "PBS/s", "cores": cores, "kind": "port", "fft": "scalar radix-2",
                   "sample": f"{n} keyswitch+bootstrap of the same parameter set in {dt:.1f} s, synthetic BSK/KSK; {CPU_KIND_NOTE}"}
        bytes_io = int(host_in.numel() * 8)
        traffic = load_traffic()
        line = {
            "metric": "PBS/s (device-timed)", "value": value, "unit": "PBS/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(params),
            "decrypted_output_correct": correct,
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp64_peak if fp64_peak else None,
                         "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                         "traffic_note": None if traffic is None else traffic.get("note"),
                         "kernel": "k_pbs_pipe", "launches_timed": n_pbs_launch, "mean_launch_ms": pbs_ms / max(n_pbs_launch, 1),
                         "ciphertexts_per_launch": pbs_cts / max(n_pbs_launch, 1),
                         "flops_per_pbs": fl, "share_of_step": pbs_ms / ms if ms else None,
                         "peak_source": "DFMA micro-benchmark run in this process (MEASURED_PEAKS.json has no fp64 entry)"},
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": "PBS/s", "h2d_bytes_per_step": bytes_io, "d2h_bytes_per_step": bytes_io,
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": launches,
            "clocks": clocks,
        }
        if configs is not None:
            line["configs"] = configs
        if strong is not None:
            line["strong_scaling"] = strong
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the per-config / strong-scaling blocks")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
