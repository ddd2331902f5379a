"""Stall samples of one kernel by source line: joins an ncu source-page CSV (SASS rows) with the line
table of `nvdisasm -g -c <cubin>`.  usage: ncu_line_samples.py src.csv dis.txt kernel-substring [lo hi (exclude range)]"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
body = [dict(zip(hdr, r)) for r in rows[h + 1:] if len(r) == len(hdr)]
base = int(body[0]["Address"], 16)
line_of = {}; cur = None; infn = False
for ln in open(sys.argv[2]):
    if ln.startswith(".text.") or re.match(r"\s*\.section\s+\.text\.", ln):
        infn = sys.argv[3] in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2)))
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/", ln)
    if m and infn: line_of[int(m.group(1), 16)] = cur
ex = (int(sys.argv[4], 16), int(sys.argv[5], 16)) if len(sys.argv) > 5 else None
cnt = collections.Counter(); reasons = collections.defaultdict(collections.Counter); tot = 0
stall_cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
for r in body:
    a = int(r["Address"], 16) - base
    s = float(r["# Samples"] or 0); tot += s
    if ex and ex[0] <= a <= ex[1]: continue
    k = line_of.get(a)
    cnt[k] += s
    for c in stall_cols: reasons[k][c] += float(r[c] or 0)
print("total samples", tot, "shown", sum(cnt.values()))
for k, v in cnt.most_common(40):
    rs = ", ".join(f"{c[6:]}={int(x)}" for c, x in reasons[k].most_common(3))
    print(f"{100*v/tot:5.2f}%  {k}   {rs}")
