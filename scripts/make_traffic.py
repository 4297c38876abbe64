"""profiles/traffic.json from an ncu summary CSV (scripts/ncu_summary.py): the per-launch figures of k_weval that bench.py's roofline
object quotes, tied to the kernel sources by their hash.
usage: make_traffic.py summary.csv workload source-note [out.json]"""
import csv, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_hash

def num(s):
    return float(s.split()[0])

def mbytes(s):
    v, u = s.split()[0], s.split()[1] if len(s.split()) > 1 else "byte"
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

path, name, note = sys.argv[1], sys.argv[2], sys.argv[3]
out = sys.argv[4] if len(sys.argv) > 4 else os.path.join(ROOT, "profiles", "traffic.json")
rows = [r for r in csv.DictReader(open(path)) if "k_weval" in r["Kernel Name"]]
r = max(rows, key=lambda r: num(r["smsp__inst_executed.sum"]))      # the full window among the captured launches
P = "sm__inst_executed_pipe_%s.avg.pct_of_peak_sustained_active"
entry = dict(
    kernel=r["Kernel Name"].split("(")[0].replace("void ", "") + ", one full window of 64 proposals per chain",
    csrc_sha16=csrc_hash(),
    us_per_launch_under_ncu=num(r["gpu__time_duration.sum"]),
    registers=int(num(r["launch__registers_per_thread"])),
    dram_bytes_per_launch=int(mbytes(r["dram__bytes_read.sum"]) + mbytes(r["dram__bytes_write.sum"])),
    inst_per_cycle_per_sm=num(r["sm__inst_executed.avg.per_cycle_elapsed"]),
    issue_active_pct=num(r["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
    warp_instructions_per_launch=int(num(r["smsp__inst_executed.sum"])),
    l1_hit_pct=num(r["l1tex__t_sector_hit_rate.pct"]),
    l1_data_pipe_pct=num(r["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]),
    pipes_pct=dict((p, num(r[P % p])) for p in ("lsu", "xu", "alu", "fp64", "fma")),
    source=note)
try:
    cur = json.load(open(out))
except Exception:
    cur = {}
cur[name] = entry
json.dump(cur, open(out, "w"), indent=1)
print(json.dumps(entry, indent=1))
