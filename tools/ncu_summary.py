"""Summaries of ncu output for profiles/ (read here, from what came back in gpurun_out/).

    python tools/ncu_summary.py launches <launch-list.csv>            # per-kernel launch count / total / share (markdown)
    python tools/ncu_summary.py full <capture.ncu-rep> [<traffic.json>] # per-launch table of a --set full capture; optionally
                                                                       # records the conv DRAM traffic per forward in traffic.json
"""
import collections
import csv
import io
import json
import subprocess
import sys


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.OrderedDict()
    for r in rows:
        if "Kernel Name" in r:
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        u = d.get("Metric Unit", "")
        v = v / 1000.0 if u in ("ns", "nsecond") else (v * 1000.0 if u in ("ms", "msecond") else v)
        k = d["Kernel Name"].split("(")[0][:70]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("| kernel | launches | total us | mean us | share |\n|---|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {v[1] / v[0]:.2f} | {100 * v[1] / tot:.1f} % |")
    print(f"\ntotal {tot:.1f} us over {sum(v[0] for v in agg.values())} launches")


METRICS = {
    "gpu__time_duration.sum": "time us",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor pipe active %",
    "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active": "tensor inst %",
    "dram__bytes_read.sum": "dram read MB",
    "dram__bytes_write.sum": "dram write MB",
    "lts__t_sector_hit_rate.pct": "L2 hit %",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "smem LSU wavefronts %",
    "launch__registers_per_thread": "regs",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "SM throughput %",
}


def full(rep, traffic_json=None):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    units = rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    cols = [(m, n) for m, n in METRICS.items() if m in idx]
    tensor_alt = [h for h in hdr if "pipe_tensor" in h and "pct" in h]
    print("| # | kernel | " + " | ".join(n for _, n in cols) + " |\n|---|---|" + "---|" * len(cols))
    total_dram = 0.0
    for k, r in enumerate(rows[2:]):
        if len(r) < len(hdr):
            continue
        name = r[idx["Kernel Name"]].split("(")[0][:48]
        vals = []
        for m, n in cols:
            try:
                v = float(r[idx[m]].replace(",", ""))
            except ValueError:
                vals.append(r[idx[m]])
                continue
            u = units[idx[m]]
            if "bytes" in m:
                v = v / (1e6 if u in ("byte", "bytes", "B") else 1e3 if u.startswith("K") else 1.0 if u.startswith("M") else 1e-3)
                if "conv3x3_halo" in name:
                    total_dram += v
            if "time" in m:
                v = v / 1000.0 if u in ("ns", "nsecond") else v
            vals.append(f"{v:.1f}")
        print(f"| {k} | `{name}` | " + " | ".join(vals) + " |")
    print(f"\nDRAM traffic of the conv3x3_halo launches in this capture: {total_dram:.1f} MB")
    if tensor_alt:
        print("tensor-pipe metrics available:", ", ".join(tensor_alt[:6]))
    return total_dram


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
