"""Print the headline counters and the top stall instructions of one kernel of an .ncu-rep (read on the CPU box)."""
import csv, subprocess, sys
rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, unit = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'sm__cycles_elapsed.max',
        'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'launch__occupancy_limit_registers']
for v in rows[2:]:
    print("=" * 100)
    for i, h in enumerate(hdr):
        if h in want or (h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')):
            try:
                if h.startswith('smsp__average') and float(v[i]) < 0.05:
                    continue
            except ValueError:
                pass
            print("%-90s %-12s %s" % (h, unit[i], v[i]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
h = None
data = []
for r in rows:
    if "Address" in r and "Source" in r:
        h = r
        ia, isrc, iall, iex = h.index("Address"), h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        continue
    if h is None or len(r) <= max(ia, isrc, iall, iex):
        continue
    try:
        data.append((int(r[iall] or 0), r[ia], r[isrc], int(r[iex] or 0)))
    except ValueError:
        pass
tot = sum(d[0] for d in data) or 1
print("total samples", tot, "warp instructions", sum(d[3] for d in data))
for d in sorted(sorted(data, key=lambda d: -d[0])[:ntop], key=lambda d: int(d[1], 16)):
    print("%6d %5.1f%% %s  %-72s %d" % (d[0], 100 * d[0] / tot, d[1][-5:], d[2][:72], d[3]))
