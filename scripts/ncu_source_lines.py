"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export per CUDA source line.
usage: ncu_source_lines.py export.csv [top_n] [kernel_index]"""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40; which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
rows = csv.reader(open(path))
fpath = None; hdr = None; fn = None
seen_fn = []
agg = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name":
        fn = r[1]
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    try: line = int(r[0])
    except ValueError: continue
    i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); i_t = hdr.index("Thread Instructions Executed")
    key = (fpath.split("/")[-1], line)
    a = agg[key]
    try:
        a[0] += int(r[i_s]); a[1] += int(r[i_i]); a[2] += int(r[i_t])
    except ValueError:
        pass
    a[3] = r[1]
tot_s = sum(a[0] for a in agg.values()); tot_i = sum(a[1] for a in agg.values())
print("total samples", tot_s, "warp-instr", tot_i)
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  lanes %4.1f  %s:%d  %s" % (100.0 * a[0] / max(tot_s, 1), 100.0 * a[1] / max(tot_i, 1), a[2] / max(a[1], 1), key[0], key[1], a[3].strip()[:110]))
