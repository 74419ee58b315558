"""profiles/<round>_kernels_ncu.md from the .ncu-rep files written by tools/profile_all.sh (read here, no GPU needed)."""
import csv
import glob
import os
import subprocess
import sys

WANT = [("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occupancy %"), ("launch__registers_per_thread", "regs"),
        ("launch__grid_size", "grid"), ("launch__block_size", "block"), ("smsp__inst_executed.sum", "warp inst"),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma %"), ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu %"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
        ("lts__t_sector_hit_rate.pct", "L2 hit %"), ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts")]


def one(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return None
    H, U, V = rows[0], rows[1], rows[2]
    d = {"kernel": V[H.index("Kernel Name")][:90]}
    for m, label in WANT:
        if m in H:
            i = H.index(m)
            d[label] = f"{V[i]} {U[i]}".strip()
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = next((r for r in rows if r and r[0] == "Address"), None)
    if hdr:
        cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        agg = {hdr[i][6:]: 0 for i in cols}
        for r in rows[rows.index(hdr) + 1:]:
            if len(r) == len(hdr) and r[0] != "Address":      # a report with several launches repeats the header per function
                for i in cols:
                    agg[hdr[i][6:]] += int(r[i] or 0) if (r[i] or "0").lstrip("-").isdigit() else 0
        tot = max(1, sum(agg.values()))
        d["top stalls"] = ", ".join(f"{k} {100 * v / tot:.0f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:4])
    return d


def main(rnd="r1", outdir="gpurun_out"):
    reps = sorted(glob.glob(os.path.join(outdir, f"prof_{rnd}_*.ncu-rep")))
    lines = [f"# ncu --set full --clock-control none, one launch per kernel ({rnd}; tools/profile_all.sh + tools/ncu_table.py)", ""]
    for rep in reps:
        d = one(rep)
        if not d:
            continue
        lines.append(f"## {os.path.basename(rep)[len('prof_' + rnd + '_'):-8]}")
        lines.append(f"`{d.pop('kernel')}`")
        lines.append("")
        for k, v in d.items():
            lines.append(f"- {k}: {v}")
        lines.append("")
    open(os.path.join("profiles", f"{rnd}_kernels_ncu.md"), "w").write("\n".join(lines))
    print("\n".join(lines))


if __name__ == "__main__":
    main(*sys.argv[1:])
