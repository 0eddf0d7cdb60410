"""Summarise one kernel of an .ncu-rep (ncu --set full) as `metric [unit] = value` lines, the
format of profiles/r01_ncu_*.txt:   python scripts/ncu_summary.py REPORT.ncu-rep "title" > profiles/x.txt"""
import csv, io, re, subprocess, sys

KEEP = re.compile(
    r"^(gpu__time_duration\.sum|dram__bytes_(read|write)\.sum|launch__(grid_size|registers_per_thread|"
    r"waves_per_multiprocessor|block_size)|lts__t_sector_hit_rate\.pct|lts__t_bytes\.sum|"
    r"sm__inst_executed_pipe_(alu|fma|fp64|lsu|xu|uniform|tc|tmem)\.avg\.pct_of_peak_sustained_active|"
    r"sm__pipe_(tc|tensor|tensor_subpipe_hmma|tensor_subpipe_dmma|shared|tma)_cycles_active\.avg\.pct_of_peak_sustained_active|"
    r"sm__throughput\.avg\.pct_of_peak_sustained_elapsed|sm__warps_active\.avg\.pct_of_peak_sustained_active|"
    r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio|smsp__inst_executed\.sum|"
    r"smsp__issue_active\.avg\.pct_of_peak_sustained_active|l1tex__data_pipe_tc_wavefronts_mem_shared\.sum|"
    r"smsp__sass_inst_executed_op_tmem_(ldt|stt)\.sum|sm__cycles_active\.avg|sm__cycles_elapsed\.max)$")


def main():
    rep, title = sys.argv[1], sys.argv[2]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    names, units, vals = rows[0], rows[1], rows[2]
    print(title)
    for n, u, v in sorted(zip(names, units, vals)):
        if KEEP.match(n):
            print("%s [%s] = %s" % (n, u, v))


if __name__ == "__main__":
    main()
