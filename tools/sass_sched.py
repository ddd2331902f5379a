"""Static issue model of the innermost shuffle loop of a kernel: decodes the scheduling control word
of every sm_100 SASS instruction (stall count, yield, write/read barrier, wait mask) and prints the
loop body with them, plus the sum of stall counts (= issue time of a lone warp if no barrier waits).
usage: python tools/sass_sched.py <object> <kernel-substring> [--dump]"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
dump = "--dump" in sys.argv
sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
cur, kernels = None, {}
lines = sass.splitlines()
i = 0
while i < len(lines):
    line = lines[i]
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1); kernels[cur] = []
        i += 1; continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);\s+/\* (0x[0-9a-f]+) \*/", line)
    if m and cur:
        m2 = re.match(r"\s+/\* (0x[0-9a-f]+) \*/", lines[i + 1])
        hi = int(m2.group(1), 16)
        ctrl = hi >> 41
        stall = ctrl & 0xf; yld = (ctrl >> 4) & 1; wbar = (ctrl >> 5) & 7; rbar = (ctrl >> 8) & 7
        wait = (ctrl >> 11) & 0x3f
        kernels[cur].append((int(m.group(1), 16), m.group(2).strip(), stall, yld, wbar, rbar, wait))
        i += 2; continue
    i += 1
for name, ins in kernels.items():
    if pat not in name:
        continue
    addr = {a: k for k, (a, *_) in enumerate(ins)}
    loops = []
    for k, (a, t, *_) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m and "BRA.DIV" not in t:
            tgt = int(m.group(1), 16)
            if tgt < a and tgt in addr:
                loops.append((addr[tgt], k))
    print(name[:100])
    for (b, e) in loops:
        body = ins[b:e + 1]
        if not any("SHFL" in t for _, t, *_ in body):
            continue
        if any(b <= b2 and e2 <= e and (b2, e2) != (b, e) for (b2, e2) in loops if any("SHFL" in x[1] for x in ins[b2:e2+1])):
            continue
        tot = sum(x[2] for x in body)
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", x[1]).split()[0].split(".")[0] for x in body)
        print(f"  loop {hex(body[0][0])}-{hex(body[-1][0])}: {len(body)} instr, sum of stall counts {tot}", dict(ops.most_common()))
        if dump:
            for (a, t, stall, yld, wbar, rbar, wait) in body:
                wb = "-" if wbar == 7 else str(wbar); rb = "-" if rbar == 7 else str(rbar)
                print(f"    {a:05x} s{stall:<2} {'Y' if yld else ' '} w{wb} r{rb} wait{wait:06b}  {t}")
