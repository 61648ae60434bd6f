"""Developer tool: dynamic instruction counts per source line, split by opcode, for chosen opcodes (ncu source page)."""
import csv, io, subprocess, sys, collections, re
rep = sys.argv[1]; recs = float(sys.argv[2]); want = sys.argv[3].split(","); top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = None; cur_line = None; cur_src = ""
agg = collections.Counter(); srcs = {}
for r in csv.reader(io.StringIO(cs)):
    if len(r) == 2 and r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0].isdigit(): cur_line = int(r[0]); cur_src = r[1].strip()[:90]; continue
    if len(r) > 10 and r[0] == "" and r[2].startswith("0x"):
        ins = re.sub(r"^@!?U?P\d+\s+", "", r[3].strip())
        op = ins.split()[0].split(".")[0]
        if op in want:
            try: n = int(r[7] or 0)
            except ValueError: n = 0
            agg[(cur_file, cur_line, op)] += n; srcs[(cur_file, cur_line)] = cur_src
tot = collections.Counter()
for (f, l, op), n in agg.items(): tot[op] += n
print("totals/rec:", {o: round(n / recs, 2) for o, n in tot.items()})
for (f, l, op), n in agg.most_common(top):
    print(f"{n / recs:6.2f}/rec {op:6s} {f}:{l}: {srcs[(f, l)]}")
