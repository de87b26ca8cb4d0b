"""Summarise an ncu capture into tracked files under profiles/.
usage: python scratch/make_profiles.py TAG REPORT.ncu-rep LAUNCHES.csv
writes profiles/TAG_ncu_summary.json, profiles/TAG_<kernel>_per_line.txt, profiles/TAG_launches.csv,
       profiles/render_kernel_traffic.json (dram bytes per launch per kernel, read by bench.py)"""
import csv, io, json, re, subprocess, sys
from pathlib import Path
tag, rep, launches = sys.argv[1:4]
opts = set(sys.argv[4:])     # "notraffic": keep render_kernel_traffic.json; "nolines": no per-line files; "merge": add to an existing summary
root = Path(__file__).resolve().parent.parent
prof = root / "profiles"
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
           "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
           "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
           "launch__occupancy_limit_shared_mem", "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, data = rows[0], rows[1], rows[2:]
kcol = hdr.index("Kernel Name")
summary, traffic = {}, {}
for r in data:
    name = re.sub(r"^void |\(.*$|<unnamed>::|<.*$", "", r[kcol])
    d = {}
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            d[m] = {"value": r[i], "unit": units[i]}
    summary.setdefault(name, []).append(d)
    def mb(m):
        i = hdr.index(m); v = float(r[i]); u = units[i].lower()
        return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
    traffic[name] = {"dram_bytes_per_launch": mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum"),
                     "source": f"profiles/{tag}_ncu_summary.json (ncu --set full --clock-control none, one launch)"}
sp = prof / f"{tag}_ncu_summary.json"
if "merge" in opts and sp.exists():
    old = json.loads(sp.read_text())
    old.update(summary)
    summary = old
sp.write_text(json.dumps(summary, indent=1))
if "notraffic" not in opts:
    (prof / "render_kernel_traffic.json").write_text(json.dumps(traffic, indent=1))
# per-line
for li, r in enumerate(data if "nolines" not in opts else []):
    name = re.sub(r"^void |\(.*$|<unnamed>::|<.*$", "", r[kcol])
    txt = subprocess.run([sys.executable, str(root / "scratch" / "ncu_lines.py"), rep, str(li), "45"], capture_output=True, text=True).stdout
    mem = subprocess.run([sys.executable, str(root / "scratch" / "ncu_mem.py"), rep, str(li), "20"], capture_output=True, text=True).stdout
    (prof / f"{tag}_{name}_per_line.txt").write_text("== instructions / stall samples per source line ==\n" + txt + "\n== memory traffic per source line ==\n" + mem)
# launch list: keep our kernels only
if launches != "-":
    keep = [l for l in open(launches) if not l.startswith("==")]
    (prof / f"{tag}_launches.csv").write_text("".join(keep))
print("wrote", sorted(p.name for p in prof.glob(tag + "*")), "render_kernel_traffic.json")
