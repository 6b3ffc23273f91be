"""Summarise an .ncu-rep (ncu --set full) into a small text table for profiles/: one row per captured launch."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("Grid Size", "grid"), ("Block Size", "block"), ("gpu__time_duration.sum", "time"),
        ("dram__bytes_read.sum", "dram_rd"), ("dram__bytes_write.sum", "dram_wr"), ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram_%"),
        ("l1tex__m_xbar2l1tex_read_bytes.sum", "l2_to_sm"), ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_%"),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue_%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ_%"), ("launch__registers_per_thread", "regs"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall_barrier"), ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall_long_sb"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall_lg_throttle"), ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall_math")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print("# " + rep)
for r in rows[2:]:
    parts = []
    for i, n in idx:
        v = r[i]
        if n == "kernel":
            v = v.split("(")[0].replace("void ", "")
        else:
            try:
                v = f"{float(v):.4g}"
            except ValueError:
                pass
            if units[i] and n not in ("grid", "block"):
                v += " " + units[i]
        parts.append(f"{n}={v}")
    print("  ".join(parts))
