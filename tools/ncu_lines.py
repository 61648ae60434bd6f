"""Summarise an .ncu-rep: headline metrics + instructions per source line (developer tool)."""
import csv, subprocess, sys, io
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__occupancy_limit", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__grid_size", "sm__cycles_elapsed.max", "lts__t_sectors_srcunit_tex_op_read.sum",
        "smsp__average_warp_latency_per_inst_issued"]
for h, u, v in zip(hdr, units, vals):
    if any(h.startswith(w) for w in want) and "pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", ""):
        print(f"{h} [{u}] = {v}")
st = []
for h, v in zip(hdr, vals):
    if "smsp__average_warps_issue_stalled" in h and "per_issue_active" in h:
        try: st.append((float(v), h.split("stalled_")[1].split("_per_issue")[0]))
        except ValueError: pass
print("stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in sorted(st, reverse=True)[:8]))
cs = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; out = []
def I(x):
    try: return int(x)
    except ValueError: return 0
for r in csv.reader(io.StringIO(cs)):
    if len(r) == 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) > 10 and r[0].isdigit():
        out.append((I(r[7]), I(r[8]), I(r[6]), cur, int(r[0]), r[1].strip()[:100]))
tot = sum(o[0] for o in out)
print("total warp-inst", tot)
out.sort(reverse=True)
acc = 0
for inst, tinst, samp, f, ln, src in out[:top]:
    acc += inst
    print(f"{inst/tot*100:5.1f}% cum {acc/tot*100:5.1f}% thr {tinst/max(1,inst):4.1f} smp {samp:5d} {f}:{ln}: {src}")
