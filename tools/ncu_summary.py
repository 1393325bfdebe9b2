"""Summarise an .ncu-rep (raw page + hot SASS) -> text.  usage: ncu_summary.py rep.ncu-rep tokens_per_launch [out.txt]"""
import csv, subprocess, sys, io
rep, ntok = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sector_hit_rate.pct','lts__t_bytes.sum','lts__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread','launch__shared_mem_per_block_dynamic','smsp__inst_executed.sum','launch__grid_size','launch__block_size',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__warps_eligible.avg.per_cycle_active','sm__cycles_elapsed.max']
out = []
d = dict(zip(hdr, vals))
for h, u, v in zip(hdr, units, vals):
    if h in keep or h.startswith('smsp__average_warps_issue_stalled'):
        out.append(f"{h} [{u}] = {v}")
try:
    inst = float(d['smsp__inst_executed.sum']); out.append(f"warp-instructions per token = {inst/ntok:.1f}")
    rd, wr = float(d['dram__bytes_read.sum']), float(d['dram__bytes_write.sum'])
    ur, uw = units[hdr.index('dram__bytes_read.sum')], units[hdr.index('dram__bytes_write.sum')]
    sc = {'Mbyte':1e6,'Gbyte':1e9,'Kbyte':1e3,'byte':1}
    tot = rd*sc[ur] + wr*sc[uw]; out.append(f"dram bytes per launch = {tot:.4g}  ({tot/ntok:.1f} B/token)")
except Exception as ex:
    out.append(f"(derived metrics failed: {ex})")
txt = "\n".join(out)
print(txt)
if len(sys.argv) > 3:
    open(sys.argv[3], "w").write(txt + "\n")
