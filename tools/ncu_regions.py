"""Developer tool: warp-instructions per record by code region, from the per-line table tools/ncu_mix.py writes (argv[4])."""
import os, sys, re
# regions for the v6 file layout: find by function names via line ranges computed from the source
src_k = open(os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/pss-bam_b200/csrc/pss_kernels.cuh").read().splitlines()
src_r = open(os.path.dirname(os.path.dirname(os.path.abspath(__file__))) + "/pss-bam_b200/csrc/pss_record.h").read().splitlines()
def find(src, pat):
    for i, l in enumerate(src, 1):
        if pat in l: return i
    raise KeyError(pat)
K = [("lop3/imad/classify", find(src_k, "template <int LUT>"), find(src_k, "__device__ __forceinline__ void log_outcome") - 1),
     ("contig", find(src_k, "// k-th 4-byte word of the field"), find(src_k, "// A record the tile cannot hold") - 1),
     ("tally_rows", find(src_k, "// ballot of \"(word & mask) != 0\""), find(src_k, "__device__ __forceinline__ void flush_acc") - 2),
     ("flush", find(src_k, "__device__ __forceinline__ void flush_acc") - 1, find(src_k, "// building blocks of the tally kernel") - 2),
     ("list_generic", find(src_k, "// Exact newline listing"), find(src_k, "// One warp-load of records") - 1),
     ("batch_glue", find(src_k, "// One warp-load of records"), find(src_k, "// tally kernel.") - 2),
     ("k_stage", find(src_k, "// ---- stage ----"), find(src_k, "// ---- pass A") - 1),
     ("k_passA", find(src_k, "// ---- pass A"), find(src_k, "// ---- pass B") - 1),
     ("k_passB", find(src_k, "// ---- pass B"), find(src_k, "// ---- records ----") - 1),
     ("k_next", find(src_k, "// ---- where the next tile begins") - 2, find(src_k, "// ---- records ----") - 1),
     ("k_records", find(src_k, "// ---- records ----"), find(src_k, "flush_acc<NACC>(acc, rows, lane, S.sh.table);\n".strip()) - 1),
     ("accessors", find(src_k, "struct SmemAt {"), find(src_k, "template <int LUT>") - 1),
     ("ptx", 20, 60)]
R = [("numbers", find(src_r, "template <class B>\nPSS_HD uint32_t word_at".split("\n")[1]) - 1, find(src_r, "PSS_HD int split_fast") - 2),
     ("split_fast", find(src_r, "PSS_HD int split_fast") - 1, find(src_r, "// genome access") - 2),
     ("load_window", find(src_r, "PSS_HD void load_window"), find(src_r, "// the byte behind an \"other\" symbol") - 1),
     ("ctx_member", find(src_r, "// the byte behind an \"other\" symbol"), find(src_r, "// CIGAR must be the bytes of") - 1),
     ("cigar", find(src_r, "// CIGAR must be the bytes of"), find(src_r, "// read bases -> 2-bit codes") - 2),
     ("read_codes", find(src_r, "// read bases -> 2-bit codes") - 1, find(src_r, "// What one record contributes") - 1),
     ("pss_record", find(src_r, "// What one record contributes"), find(src_r, "// fragkon.c:122-216 process_aln.") - 2),
     ("bithelpers", find(src_r, "// small bit helpers with host twins"), find(src_r, "// outcomes and options") - 2),
     ("scan11", find(src_r, "PSS_HD_NOINLINE int scan11") - 1, find(src_r, "// split_fast: a record") - 2)]
agg = {}; smp = {}; thr = {}
for l in open(sys.argv[1]):
    loc, inst, per, t, s, src = l.rstrip("\n").split("\t", 5)
    f, ln = loc.rsplit(":", 1); ln = int(ln)
    tab = K if "kernels" in f else (R if "record" in f else [])
    name = f"{f}:other"
    for nm, a, b in tab:
        if a <= ln <= b: name = nm
    agg[name] = agg.get(name, 0) + float(per); smp[name] = smp.get(name, 0) + int(s)
    thr[name] = thr.get(name, 0) + float(per) * float(t)
tot = sum(agg.values()); ts = sum(smp.values())
for n, v in sorted(agg.items(), key=lambda x: -x[1]):
    if v > 0: print(f"{n:28s} {v:7.2f}/rec {v / tot * 100:5.1f}%  thr {thr[n] / v:4.1f}  samples {smp[n] / ts * 100:5.1f}%")
print("total", tot)
