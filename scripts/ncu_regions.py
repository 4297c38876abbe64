"""Per-kernel, per-source-line aggregation of an `ncu --page source --csv --print-source cuda,sass` export.
usage: ncu_regions.py export.csv kernel_name [top_n]"""
import csv, sys, collections
path, kname_want = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rows = csv.reader(open(path))
fpath = hdr = fn = None
A = collections.defaultdict(lambda: [0, 0, 0, ""])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fpath = r[1]; continue
    if r[0] == "Function Name": fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or r[0] == "": continue
    try: line = int(r[0])
    except ValueError: continue
    kname = fn.split('(')[0].split('<')[0].replace('void ', '')
    if kname != kname_want: continue
    i_s = hdr.index("# Samples"); i_i = hdr.index("Instructions Executed"); i_t = hdr.index("Thread Instructions Executed")
    a = A[(fpath.split('/')[-1], line)]
    try: a[0] += int(r[i_s]); a[1] += int(r[i_i]); a[2] += int(r[i_t])
    except ValueError: pass
    a[3] = r[1]
ts = sum(a[0] for a in A.values()); ti = sum(a[1] for a in A.values())
print(kname_want, "samples", ts, "warp-instr", ti)
F = collections.defaultdict(lambda: [0, 0])
for k, a in A.items(): F[k[0]][0] += a[0]; F[k[0]][1] += a[1]
for k, v in sorted(F.items(), key=lambda kv: -kv[1][0])[:8]: print("  file %5.1f%% smp %5.1f%% inst  %s" % (100 * v[0] / ts, 100 * v[1] / ti, k))
for key, a in sorted(A.items(), key=lambda kv: -kv[1][0])[:top]:
    print("%5.1f%% smp %5.1f%% inst  lanes %4.1f  %s:%d  %s" % (100.0 * a[0] / ts, 100.0 * a[1] / ti, a[2] / max(a[1], 1), key[0], key[1], a[3].strip()[:100]))
