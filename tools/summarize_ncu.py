"""Turns gpurun_out/launches.csv and gpurun_out/prof_gemm.ncu-rep into the tracked summaries under profiles/.
Usage: python tools/summarize_ncu.py <round tag, e.g. r01> ; runs on the CPU box (ncu -i needs no GPU)."""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")


def launch_list(tag):
    rows = [r for r in csv.reader(open(os.path.join(OUT, "launches.csv"))) if len(r) > 10]
    hdr = rows[0]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ik].split("(")[0]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    lines = ["| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"| `{k}` | {n} | {t / 1e6:.3f} | {t / tot:.3f} |")
    return lines, {k: {"launches": n, "ms": t / 1e6, "share": t / tot} for k, (n, t) in agg.items()}


WANT = {
    "gpu__time_duration.sum": "duration",
    "dram__bytes_read.sum": "dram read",
    "dram__bytes_write.sum": "dram write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram % of peak",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor pipe active %",
    "sm__inst_executed_pipe_tensor_op_hmma.avg.pct_of_peak_sustained_active": "tensor hmma %",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor hmma pipe active %",
    "l1tex__m_xbar2l1tex_read_bytes.sum": "L2 -> SM bytes",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "L2 throughput %",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "SM throughput %",
    "lts__t_sector_hit_rate.pct": "L2 hit rate %",
    "launch__registers_per_thread": "registers/thread",
    "launch__grid_size": "grid",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps active %",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem wavefronts",
}


def full_capture():
    rep = os.path.join(OUT, "prof_gemm.ncu-rep")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = {}
    for i, h in enumerate(hdr):
        for key in WANT:
            if h == key or (key.startswith("sm__pipe_tensor") and h.startswith("sm__pipe_tensor_cycles_active") and "pct" in h):
                cols.setdefault(h, i)
    names = [r[hdr.index("Kernel Name")] for r in data]
    lines = ["| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |", "|---|---|" + "---|" * len(data)]
    out = []
    for h, i in cols.items():
        lines.append(f"| {h} | {units[i]} | " + " | ".join(r[i] for r in data) + " |")
    def val(metric, r):
        i = hdr.index(metric)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        return v * {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1}.get(u, 1)
    traffic = [val("dram__bytes_read.sum", r) + val("dram__bytes_write.sum", r) for r in data]
    return lines, names, traffic


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    dst = os.path.join(ROOT, "profiles", tag)
    os.makedirs(dst, exist_ok=True)
    l_lines, l_json = launch_list(tag)
    f_lines, names, traffic = full_capture()
    with open(os.path.join(dst, "ncu_summary.md"), "w") as f:
        f.write(f"# ncu summary ({tag})\n\nCommand: `python bench.py --channels 1 --minutes 10 --steps 1 --warmup 1 --no-cpu-baseline --train-steps 0` "
                "(one 10-minute channel = 60 000 windows, 2 chunks x 19 conv launches per pass).\n\n"
                "## Launch list (`--metrics gpu__time_duration.sum --clock-control none`, cold-cache, serialised: compare SHARES)\n\n")
        f.write("\n".join(l_lines) + "\n\n## Full capture of the block1 conv launches (`--set full`, launches = block1.0.conv1, "
                "block1.0.conv2, block1.1.conv1, block1.1.conv2 of one chunk)\n\n")
        f.write("\n".join(f_lines) + "\n\n")
        f.write("DRAM traffic per launch (read + write): " + ", ".join(f"{t / 1e9:.3f} GB" for t in traffic) + "\n")
    latest = {"round": tag, "launch_list": l_json, "gemm_dram_bytes_per_launch": traffic[-1],
              "gemm_dram_bytes_per_launch_all": traffic, "gemm_capture": "block1.1.conv2, chunk of 32768 window starts"}
    with open(os.path.join(ROOT, "profiles", "latest.json"), "w") as f:
        json.dump(latest, f, indent=1)
    print(open(os.path.join(dst, "ncu_summary.md")).read())


if __name__ == "__main__":
    main()
