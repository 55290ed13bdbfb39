"""FP64 pipe cost of a SASS address range under the measured operand-bandwidth model (tools/fp64_operands.cu):
a warp-wide FP64 instruction occupies its sub-partition's FP64 path for max(2, R) cycles, R = distinct 64-bit
REGISTER source operands that are not served by the operand reuse cache (uniform registers, constants and
immediates are free).
python tools/sass_fp64_cost.py <file.o> <kernel-substring> <lo-hex> <hi-hex> [skip-lo skip-hi ...]"""
import re, subprocess, sys
path, pat, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3], 16), int(sys.argv[4], 16)
skips = [(int(sys.argv[i], 16), int(sys.argv[i + 1], 16)) for i in range(5, len(sys.argv) - 1, 2)]
txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout.splitlines()
cur = None; ins = []
for ln in txt:
    m = re.search(r"Function : (\S+)", ln)
    if m: cur = m.group(1); continue
    if cur is None or pat not in cur: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
prev = None; tot = 0; n = 0; hist = {}
for a, t in ins:
    if not (lo <= a <= hi) or any(s0 <= a <= s1 for s0, s1 in skips): prev = None if not (lo <= a <= hi) else prev; continue
    t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t2.split()[0].split(".")[0]
    if op not in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"):
        if op not in ("MOV", "IMAD", "SEL", "FSEL", "LOP3", "ISETP", "IADD3", "VIADD", "NOP", "PLOP3", "CS2R", "LEA"): prev = None   # anything else may clobber the cache model: be conservative
        continue
    ops = [o.strip() for o in t2[len(t2.split()[0]):].split(",")]
    srcs = ops[1:] if op != "DSETP" else ops[2:]
    slots = []
    for o in srcs:
        m = re.match(r"^[-|~!]*\|?(R\d+)\|?(\.reuse)?", o)
        slots.append((m.group(1), bool(m.group(2))) if m and not o.lstrip("-|").startswith("RZ") else None)
    reads = set()
    for k, s in enumerate(slots):
        if s is None: continue
        if prev is not None and k < len(prev) and prev[k] is not None and prev[k] == (s[0], True): continue
        reads.add(s[0])
    c = max(2, len(reads)); tot += c; n += 1; hist[len(reads)] = hist.get(len(reads), 0) + 1
    prev = slots
print("FP64 instructions %d, pipe cycles %d (%.2f per instruction); register reads histogram %s" % (n, tot, tot / max(n, 1), dict(sorted(hist.items()))))
