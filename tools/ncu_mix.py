"""Developer tool: dynamic opcode mix, pipe shares and per-source-line instruction counts of the first kernel in an .ncu-rep."""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]
recs = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0      # records processed by the launch: prints per-record figures
top = int(sys.argv[3]) if len(sys.argv) > 3 else 60
def page(src):
    return subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", src], capture_output=True, text=True).stdout
FMA = ("IMAD", "IDP", "FFMA", "FMUL", "FADD", "HFMA2", "HADD2", "HMUL2", "IMUL")
XU = ("POPC", "FLO", "BREV", "MUFU")
MEM = ("LDS", "STS", "LDG", "STG", "LD", "ST", "ATOMS", "ATOMG", "RED", "LDL", "STL", "LDC", "LDCU", "UBLKCP", "SYNCS", "ATOM")
CTRL = ("BRA", "BSSY", "BSYNC", "BAR", "WARPSYNC", "ENDCOLLECTIVE", "EXIT", "NOP", "CALL", "RET", "BREAK", "YIELD")
def pipe(op):
    b = op.split(".")[0]
    if b.startswith("U") and b not in ("UBLKCP",): return "uniform"
    if b in FMA: return "fma"
    if b in XU: return "xu"
    if b in MEM: return "mem"
    if b in CTRL: return "ctrl"
    if b in ("SHFL", "VOTE", "VOTEU", "MATCH", "REDUX", "S2R", "S2UR", "R2UR", "CS2R"): return "misc"
    return "alu"
ops = collections.Counter(); pipes = collections.Counter(); tot = 0; samples = collections.Counter()
for r in csv.reader(io.StringIO(page("sass"))):
    if len(r) > 6 and r[0].startswith("0x"):
        ins = r[1].strip()
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        op = ins.split()[0].rstrip(";")
        n = int(r[5] or 0); tot += n
        ops[op.split(".")[0]] += n; pipes[pipe(op)] += n; samples[pipe(op)] += int(r[4] or 0)
print(f"total warp-instructions {tot}" + (f"  = {tot / recs:.1f} per record" if recs else ""))
for p, n in pipes.most_common():
    print(f"  pipe {p:8s} {n / tot * 100:5.1f}%" + (f"  {n / recs:6.1f}/record" if recs else "") + f"   samples {samples[p]}")
print("opcodes:", ", ".join(f"{o} {n / tot * 100:.1f}%" for o, n in ops.most_common(28)))
cur = None; lines = []
for r in csv.reader(io.StringIO(page("cuda,sass"))):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0].isdigit():
        try: lines.append((int(r[7] or 0), int(r[8] or 0), int(r[6] or 0), cur, int(r[0]), r[1].strip()[:90]))
        except ValueError: pass
lt = sum(l[0] for l in lines) or 1
# per file/line-range buckets (function-ish): group by file and hundred-line block
print("per source line (top):")
for inst, tinst, smp, f, ln, src in sorted(lines, reverse=True)[:top]:
    print(f"{inst / lt * 100:5.2f}%" + (f" {inst / recs:6.2f}/rec" if recs else "") + f" thr {tinst / max(1, inst):4.1f} smp {smp:5d} {f}:{ln}: {src}")
if len(sys.argv) > 4:
    with open(sys.argv[4], "w") as fo:
        for inst, tinst, smp, f, ln, src in sorted(lines, key=lambda l: (l[3] or "", l[4])):
            if inst: fo.write(f"{f}:{ln}\t{inst}\t{inst / recs if recs else 0:.3f}\t{tinst / max(1, inst):.1f}\t{smp}\t{src}\n")
