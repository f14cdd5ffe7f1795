"""Summaries of ncu output for profiles/.
  ncu_summary.py launches <launch-list csv>            -> per-kernel launches / total ms / share
  ncu_summary.py full <raw csv> <title line>           -> section metrics + stall samples of one `--set full` capture
(raw csv = `ncu -i X.ncu-rep --page raw --csv`)"""
import csv, re, sys, collections

mode = sys.argv[1]
if mode == "launches":
    rows = list(csv.reader(l for l in open(sys.argv[2]) if l.startswith('"')))
    hdr = rows[0]
    kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if len(r) <= mv:
            continue
        name = re.sub(r"<.*", "", r[kn].replace("void ", "")).split("(")[0]
        v = float(r[mv].replace(",", ""))
        ms = v / 1e6 if r[mu] in ("ns", "nsecond") else v / 1e3 if r[mu] in ("us", "usecond") else v
        agg[name][0] += 1
        agg[name][1] += ms
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':40s} {'launches':>8s} {'total ms':>12s} {'share':>8s}")
    for k, (c, ms) in sorted(agg.items(), key=lambda t: -t[1][1]):
        print(f"{k:40s} {c:8d} {ms:12.3f} {100 * ms / tot:7.2f}%")
else:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr, units, vals = rows[0], rows[1], rows[2]
    print(sys.argv[3])
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print("kernel:", d.get("Kernel Name", ("?",))[0][:150])
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__cluster_size", "launch__registers_per_thread",
            "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.per_cycle_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
            "smsp__inst_executed.sum", "sass__inst_executed_shared_loads", "sass__inst_executed_shared_stores",
            "sass__inst_executed_global_loads", "sass__inst_executed_register_spilling", "smsp__sass_inst_executed_op_tmem_ldt.sum",
            "smsp__inst_executed_op_tma_ld.sum"]
    for w in want:
        if w in d:
            print(f"{w:96s} {d[w][0]:>20s} {d[w][1]}")
    st = {h[len("smsp__pcsamp_warps_issue_stalled_"):]: float(v) for h, v in zip(hdr, vals)
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in h and v not in ("", "n/a")}
    tot = sum(st.values())
    print("\nwarp stall samples (smsp__pcsamp_warps_issue_stalled_*):")
    for k, v in sorted(st.items(), key=lambda t: -t[1])[:12]:
        print(f"   {k:24s} {int(v):9d}  {100 * v / tot:5.1f} %")
