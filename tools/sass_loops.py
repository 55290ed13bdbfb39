"""Static look at a kernel's SASS: finds the backward branches (loops), and for each loop body prints the
instruction count, the FP64 / shuffle / select / move mix and the sum of the scheduler's stall counts
(control word bits 105..108 of each 128-bit instruction), i.e. the issue cycles ONE warp needs per trip
if nothing else interferes.
python tools/sass_loops.py <file.o|.so> <kernel-name-substring> [--min 100]"""
import re, subprocess, sys, collections

def main():
    path, pat = sys.argv[1], sys.argv[2]
    mn = int(sys.argv[sys.argv.index("--min") + 1]) if "--min" in sys.argv else 100
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True).stdout.splitlines()
    funcs, cur = {}, None
    for ln in txt:
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1); funcs[cur] = []; continue
        if cur is None: continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if m:
            funcs[cur].append([int(m.group(1), 16), m.group(2).strip(), int(m.group(3), 16), None]); continue
        m = re.match(r"\s+/\* 0x([0-9a-f]{16}) \*/", ln)
        if m and funcs[cur] and funcs[cur][-1][3] is None:
            funcs[cur][-1][3] = int(m.group(1), 16)
    for name, ins in funcs.items():
        if pat not in name: continue
        print("==", name, len(ins), "instructions,", len(ins) * 16, "bytes")
        addr2i = {a: i for i, (a, _, _, _) in enumerate(ins)}
        loops = []
        for i, (a, t, lo, hi) in enumerate(ins):
            m = re.search(r"\bBRA\S*\s+(?:\S+,\s*)?0x([0-9a-f]+)", t)
            if m:
                tgt = int(m.group(1), 16)
                if tgt <= a and tgt in addr2i: loops.append((addr2i[tgt], i))
        for (b, e) in loops:
            n = e - b + 1
            if n < mn: continue
            ops = collections.Counter(); stall = 0
            for (a, t, lo, hi) in ins[b:e + 1]:
                t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
                op = t2.split()[0].split(".")[0]
                ops[op] += 1
                if hi is not None: stall += (hi >> 41) & 0xf
            f64 = sum(v for k, v in ops.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX", "MUFU"))
            print("loop %05x..%05x: %d instr, stall-sum %d cycles, FP64-pipe %d (DFMA %d DMUL %d DADD %d DSETP %d MUFU %d), "
                  "SHFL %d, SEL/FSEL %d, MOV/IMAD.MOV %d, LDS %d, STG %d, BAR %d, SYNCS %d" % (
                      ins[b][0], ins[e][0], n, stall, f64, ops["DFMA"], ops["DMUL"], ops["DADD"], ops["DSETP"], ops["MUFU"],
                      ops["SHFL"], ops["SEL"] + ops["FSEL"], ops["MOV"] + ops["IMAD"], ops["LDS"], ops["STG"], ops["BAR"], ops["SYNCS"]))
            if "--ops" in sys.argv: print("   ", dict(ops.most_common()))

main()
