"""Turn ncu CSV exports into the small summaries kept under profiles/.
  launches:  python tools/ncu_summary.py launches <ncu --csv launch list> <out.csv> "<header comment>"
  raw:       python tools/ncu_summary.py raw <ncu -i rep --page raw --csv export> <out.json> "<source note>"
"""
import csv
import json
import re
import sys


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = name.replace("ak::", "")
    name = re.sub(r"\((?:int|bool)\)", "", name)
    return re.sub(r"\(.*$", "", name)


def launches(src, dst, note):
    rows = [l for l in open(src) if not l.startswith("==")]
    agg = {}
    for r in csv.DictReader(rows):
        if r["Metric Name"] != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3}.get(r["Metric Unit"], 1.0)
        k = (short(r["Kernel Name"]), r.get("Grid Size", ""), r.get("Block Size", ""))
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        for line in note.split("\\n"):
            f.write("# " + line + "\n")
        f.write("kernel,launches,total_us,avg_us,share,grid,block\n")
        for (k, g, b), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"\"{k}\",{n},{t:.1f},{t / n:.1f},{t / tot:.4f},\"{g}\",\"{b}\"\n")


WANT = ["Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
        "launch__shared_mem_per_block_dynamic"]


def raw(src, dst, note):
    rows = list(csv.reader(l for l in open(src) if not l.startswith("==")))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = {"_source": note, "units": {h: units[idx[h]] for h in WANT if h in idx}, "kernels": []}
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        out["kernels"].append({h: (short(r[idx[h]]) if h == "Kernel Name" else r[idx[h]]) for h in WANT if h in idx})
    json.dump(out, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
