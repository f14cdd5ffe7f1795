"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, cur = None, None
agg = {}
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Name":
        cur_file = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < 8: continue
    if r[0] != "":
        cur = (cur_file, int(r[0]), r[1].strip()[:100]); continue
    try: s = int(r[6])
    except ValueError: continue
    d = agg.setdefault(cur, {"samples": 0, "inst": 0, "stalls": {}})
    d["samples"] += s
    try: d["inst"] += int(r[7])
    except ValueError: pass
    for name in ("stall_long_sb", "stall_short_sb", "stall_wait", "stall_barrier", "stall_membar", "stall_math", "stall_lg", "stall_mio"):
        if name in hdr:
            try: d["stalls"][name] = d["stalls"].get(name, 0) + int(r[hdr.index(name)])
            except (ValueError, IndexError): pass
tot = sum(d["samples"] for d in agg.values())
print("total samples", tot)
for k, d in sorted(agg.items(), key=lambda t: -t[1]["samples"])[:top]:
    st = " ".join(f"{n[6:]}={v}" for n, v in sorted(d["stalls"].items(), key=lambda t: -t[1])[:3] if v)
    print(f"{100*d['samples']/tot:5.1f}% {k[0]}:{k[1]:<4d} inst={d['inst']:>10d} {k[2][:80]:80s} | {st}")
