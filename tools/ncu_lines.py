"""Summarises an ncu report: python tools/ncu_lines.py REPORT.ncu-rep [N] — top source lines by executed instructions,
stall samples per file, static SASS size."""
import collections, csv, subprocess, sys, io
rep = sys.argv[1]; N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
for h, u, v in zip(r[0], r[1], r[2]):
    if h in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "launch__registers_per_thread",
             "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
             "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
             "smsp__sass_inst_executed_op_local_ld.sum"):
        print(f"{h:60s} {v} {u}")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
for x in rows:
    if x and x[0] == "Line No": hdr = x; break
idx = {h: i for i, h in enumerate(hdr)}
cols = ["# Samples", "stall_no_inst", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_branch_resolving", "stall_math", "stall_selected"]
ci = [idx[c] for c in cols]
cur = None; per = collections.OrderedDict(); pf = collections.defaultdict(lambda: [0] * len(cols)); static = collections.Counter(); line = None
for x in rows:
    if not x: continue
    if x[0] == "File Path": cur = x[1].split("/")[-1]; continue
    if x[0] in ("Function Name", "Line No"): continue
    if x[0] != "" and x[2] == "-":
        line = (cur, int(x[0]))
        try: per[(cur, int(x[0]), x[1][:80])] = (int(x[7]), int(x[6]), [int(x[c]) if x[c].isdigit() else 0 for c in ci])
        except ValueError: pass
        for j, c in enumerate(ci):
            try: pf[cur][j] += int(x[c])
            except ValueError: pass
    elif x[0] == "" and x[2].startswith("0x"): static[cur] += 1
tot = sum(v[0] for v in per.values())
print("total inst %.2f G; static SASS per file: %s" % (tot / 1e9, dict(static)))
print(cols)
for k, v in pf.items(): print(" ", k, v)
for (f, l, s), (n, smp, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:N]:
    print(f"{f}:{l:4d} {n/1e6:9.1f}M {100*n/tot:5.1f}% samp {smp:7d}  {s}   ## " + " ".join(str(v) for v in st[1:]))
