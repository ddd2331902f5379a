"""Print the instruction mix of the innermost loops that contain warp shuffles (ring / intra passes).
usage: python tools/sass_loop.py <object-or-so> <kernel-substring>"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, kernels = None, {}
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", line)
    if m and cur:
        kernels[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in kernels.items():
    if pat not in name:
        continue
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and "BRA.DIV" not in t:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                loops.append((addr[tgt], i))
    print(name[:90])
    for (b, e) in loops:
        body = ins[b:e + 1]
        if not any("SHFL" in t for _, t in body):
            continue
        # innermost only
        if any(b <= b2 and e2 <= e and (b2, e2) != (b, e) for (b2, e2) in loops if any("SHFL" in t for _, t in ins[b2:e2+1])):
            continue
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in body)
        print(f"  loop {hex(ins[b][0])}-{hex(ins[e][0])}: {len(body)} instr  ", dict(ops.most_common()))
