"""Summaries of ncu outputs: launch list (csv log) and per-source-line hot spots of a .ncu-rep."""
import collections
import csv
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = [i for i, r in enumerate(rows) if r[0] == "ID"][0]
    H, data = rows[hdr], rows[hdr + 1:]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    d = collections.defaultdict(list)
    for r in data:
        d[r[ki][:70]].append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in d.values())
    for k, v in sorted(d.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k:70s} n={len(v):3d} avg={sum(v)/len(v)/1000:9.2f} us  share={100*sum(v)/tot:5.1f}%")


RAW = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
       "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
       "launch__registers_per_thread", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_shared_mem",
       "launch__occupancy_limit_registers", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
       "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
       "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    H = rows[0]
    for w in RAW:
        if w in H:
            i = H.index(w)
            print(f"{w:70s} {[r[i] for r in rows[1:]]}")
    stall = [(h, rows[2][i]) for i, h in enumerate(H) if "warp_issue_stalled" in h and h.endswith("per_warp_active.pct")]
    for h, v in sorted(stall, key=lambda t: -float(t[1] or 0))[:8]:
        print(f"  stall {h.replace('smsp__warp_issue_stalled_','').replace('_per_warp_active.pct',''):30s} {v}")


def lines(rep, top=30):
    top = int(top)
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True,
                         text=True).stdout
    f, agg = None, []
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            f = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0] != "" and len(r) > 7 and r[2] == "-":
            try:
                agg.append((int(r[7].replace(",", "")), int(r[4] or 0), f, r[0], r[1].strip()[:100]))
            except ValueError:
                pass
    tot, tots = sum(a[0] for a in agg), max(1, sum(a[1] for a in agg))
    print("total warp instructions", tot, "stall samples", tots)
    for a in sorted(agg, reverse=True)[:top]:
        print(f"{a[0]/tot*100:5.1f}% inst {a[1]/tots*100:5.1f}% stall  {a[2]}:{a[3]}  {a[4]}")


if __name__ == "__main__":
    {"launches": launches, "raw": raw, "lines": lines}[sys.argv[1]](*sys.argv[2:])
