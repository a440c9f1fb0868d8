"""Summarise an .ncu-rep (read here, no GPU): key raw metrics + hottest SASS lines."""
import csv, subprocess, sys, io

def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "sm__cycles_elapsed.avg.per_second",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio"]

def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    hdr, units, rows = raw(rep)
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("%-75s %-10s %s" % (w, units[i], " | ".join(r[i] for r in rows)))
    print("--- stall breakdown (warps per issue-active cycle), launch 0 ---")
    st = [(float(rows[0][i]), h) for i, h in enumerate(hdr) if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
    for v, h in sorted(st, reverse=True)[:8]:
        print("  %-8.3f %s" % (v, h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    # first kernel block only
    h = next(k for k, r in enumerate(rows) if r and r[0] == "Address")
    hd = rows[h]
    si, src, ie = hd.index("# Samples"), hd.index("Source"), hd.index("Instructions Executed")
    data = []
    for k, r in enumerate(rows[h + 1:]):
        if not r or r[0] in ("Kernel Name", "Address"):
            break
        data.append((int(r[si] or 0), int(r[ie] or 0), k, r[src].strip()))
    tot = sum(d[0] for d in data)
    print("--- hottest SASS lines (samples of %d; instr executed; line#) ---" % tot)
    for s_, ie_, k, t in sorted(data, reverse=True)[:top]:
        print("  %5d %5.1f%% %9d  #%-5d %s" % (s_, 100.0 * s_ / max(tot, 1), ie_, k, t))
    print("total SASS lines", len(data), "total warp-instr", sum(d[1] for d in data))

main()
