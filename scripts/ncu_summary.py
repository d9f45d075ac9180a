"""Summarises an ncu launch list (csv) and a --set full report into markdown for profiles/."""
import collections, csv, subprocess, sys

def launches(path):
    rows = list(csv.reader(l for l in open(path) if not l.startswith("==")))
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) > vi:
            agg.setdefault(r[ki].split("(")[0][:70], []).append(float(r[vi].replace(",", "")))
    tot = sum(sum(v) for v in agg.values())
    out = ["| kernel | launches | mean us | total ms | share |", "|---|---:|---:|---:|---:|"]
    for k, v in agg.items():
        out.append(f"| `{k}` | {len(v)} | {sum(v)/len(v)/1e3:.1f} | {sum(v)/1e6:.3f} | {100*sum(v)/tot:.1f}% |")
    return "\n".join(out)

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_read.sum.per_second",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_op_imma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active"]

def full(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(rows) - 2)) + " |",
           "|---|---|" + "---:|" * (len(rows) - 2)]
    ki = hdr.index("Kernel Name")
    names = [r[ki].split("(")[0] for r in rows[2:]]
    out.insert(0, "kernels: " + ", ".join(f"`{n}`" for n in names) + "\n")
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            out.append(f"| {w} | {units[i]} | " + " | ".join(r[i] for r in rows[2:]) + " |")
    return "\n".join(out)

if __name__ == "__main__":
    kind, path = sys.argv[1], sys.argv[2]
    print(launches(path) if kind == "launches" else full(path))
