"""Per-opcode stall-sample breakdown of the hottest SASS loop in an ncu source-page CSV.
usage: ncu -i rep --page source --csv > src.csv; python tools/ncu_loop_stalls.py src.csv [lo_hex hi_hex]"""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
body = [dict(zip(hdr, r)) for r in rows[h + 1:] if len(r) == len(hdr)]
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
def num(x):
    try: return float(x)
    except: return 0.0
base = int(body[0]["Address"], 16)
lo = int(sys.argv[2], 16) if len(sys.argv) > 2 else None
hi = int(sys.argv[3], 16) if len(sys.argv) > 3 else None
tot = collections.Counter(); by_op = collections.defaultdict(collections.Counter); n_inst = 0; exec_ = 0
allsamp = sum(num(r["# Samples"]) for r in body)
for r in body:
    a = int(r["Address"], 16) - base
    if lo is not None and not (lo <= a <= hi): continue
    op = re.sub(r"^@!?U?P\d+\s+", "", r["Source"]).split()[0].split(".")[0]
    n_inst += 1; exec_ += num(r["Instructions Executed"])
    for c in stall_cols:
        v = num(r[c]); tot[c] += v; by_op[op][c] += v
s = sum(tot.values())
print(f"range instr={n_inst} executed={exec_:.3e} samples={s:.0f} ({100*s/allsamp:.1f}% of kernel)")
print("by reason:", ", ".join(f"{k[6:]}={100*v/s:.1f}%" for k, v in tot.most_common(9)))
for op, c in sorted(by_op.items(), key=lambda kv: -sum(kv[1].values()))[:14]:
    t = sum(c.values())
    print(f"  {op:8s} {100*t/s:5.1f}%  " + ", ".join(f"{k[6:]}={100*v/t:.0f}%" for k, v in c.most_common(4)))
