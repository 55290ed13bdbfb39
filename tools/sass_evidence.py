"""SASS evidence file of the LM kernel (profiles/r02_sass_k_lm_solve7.txt): opcode counts over the whole kernel,
every TMA bulk copy / mbarrier instruction with two lines of context, the two sweep loops (tools/sass_loops.py)
and the FP64 operand-read cost of the fused loop (tools/sass_fp64_cost.py).
python tools/sass_evidence.py [librsdsfm.so] > profiles/r02_sass_k_lm_solve7.txt"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rs-aware-differential-sfm_b200", "librsdsfm.so")
PAT = "k_lm_solveILi7"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout.splitlines()
cur, ins = None, []
for ln in txt:
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1); continue
    if cur is None or PAT not in cur: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))

def opcode(t):
    return re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]

ops = collections.Counter(opcode(t) for _, t in ins)
print("SASS of k_lm_solve<7> in the shipped rs-aware-differential-sfm_b200/librsdsfm.so (cuobjdump -sass, sm_100a cubin)")
print("instruction counts over the whole kernel (%d instructions):" % len(ins))
for k in ("UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "MUFU", "SHFL", "REDUX", "CREDUX", "ATOMG", "LDS", "STG", "LDG", "BAR", "MEMBAR", "ELECT"):
    print("  %-8s %d" % (k, ops[k]))
print()
print("TMA bulk copies, mbarrier operations and their neighbourhood (every UBLKCP / SYNCS line with 2 lines of context):")
keep = set()
for i, (_, t) in enumerate(ins):
    if opcode(t) in ("UBLKCP", "SYNCS"):
        keep.update(range(max(0, i - 2), min(len(ins), i + 3)))
last = -2
for i in sorted(keep):
    if i != last + 1: print("  ...")
    print("  %04x  %s" % ins[i])
    last = i
sys.stdout.flush()
out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_loops.py"), lib, PAT, "--min", "300"], capture_output=True, text=True).stdout.splitlines()
print("\n".join(out[:3]))
m = re.match(r"loop ([0-9a-f]+)\.\.([0-9a-f]+):", out[2]) if len(out) > 2 else None
if m:
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sass_fp64_cost.py"), lib, PAT, m.group(1), m.group(2)], capture_output=True, text=True)
    print(r.stdout.strip().splitlines()[-1] if r.stdout.strip() else r.stderr.strip())
