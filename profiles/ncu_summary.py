"""Prints the headline rows of an `ncu --page raw --csv` export: per kernel launch the duration, executed warp
instructions, pipe utilisation, issue-slot use, stall reasons per issued instruction and DRAM traffic.
Usage: python profiles/ncu_summary.py profiles/r02m_rollout_raw.csv [...]"""
import csv
import sys

KEYS = [("ms", "gpu__time_duration.sum"), ("warp_inst", "smsp__inst_executed.sum"),
        ("issue%", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        ("fp64%", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("fma%", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("alu%", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("xu%", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed"),
        ("tensor%", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("lsu%", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed"),
        ("regs", "launch__registers_per_thread"), ("warps/SM", "sm__warps_active.avg.per_cycle_active"),
        ("dram_rd_MB", "dram__bytes_read.sum"), ("dram_wr_MB", "dram__bytes_write.sum")]
for path in sys.argv[1:]:
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        name = r[ix["Kernel Name"]]
        print(f"{path}: {name[:150]}")
        out = []
        for label, k in KEYS:
            if k in ix and r[ix[k]] != "":
                v = float(r[ix[k]].replace(",", ""))
                u = units[ix[k]]
                if label.startswith("dram"):
                    v = v * {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(u, 1.0)
                elif label == "ms":
                    v = v * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(u, 1.0)
                out.append(f"{label}={v:.4g}")
        print("   " + "  ".join(out))
        st = []
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                v = float(r[ix[h]] or 0)
                if v >= 0.1:
                    st.append((v, h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
        print("   stalls per issued instruction: " + ", ".join(f"{n} {v:.2f}" for v, n in sorted(st, reverse=True)))
