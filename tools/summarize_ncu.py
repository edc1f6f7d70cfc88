"""Turn gpurun_out/{launches.csv, prof_head.ncu-rep} into the tracked summaries under profiles/.

    python tools/summarize_ncu.py r01a      # writes profiles/r01a_launches.md, profiles/r01a_ncu_top.md
"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

KEEP = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.max",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
]


def launches(tag):
    path = os.path.join(GP, "launches.csv")
    if not os.path.exists(path):
        return
    lines = [l for l in open(path) if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        try:
            t = float(row["Metric Value"].replace(",", ""))
        except (ValueError, KeyError):
            continue
        u = row["Metric Unit"]
        t = t / 1e3 if u == "ns" else t * 1e3 if u == "ms" else t
        name = row["Kernel Name"]
        if "gemm_simt" in name:
            m = re.findall(r"ciaosr::(\w+)", name)
            name = "ciaosr::gemm_simt<" + ",".join(m[1:4]) + ">"
        name = re.sub(r"\(.*", "", re.sub(r"<.*", "", name) if "gemm_simt" not in name else name)[:90]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        tot += t
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: every kernel of one generator forward (bench.py workload), ncu launch list\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off "
                "python tools/ncu_target.py` (one step after 3 warm-up steps; cold-cache, serialised: "
                "compare SHARES, not absolutes).\n\n"
                f"total {tot / 1e3:.2f} ms over {sum(a[0] for a in agg.values())} launches\n\n"
                "| us | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {t:.1f} | {100 * t / tot:.1f}% | {n} | `{k}` |\n")
    print("wrote", f"{tag}_launches.md")


def ncu_top(tag, rep="prof_head.ncu-rep"):
    path = os.path.join(GP, rep)
    if not os.path.exists(path):
        return
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    traffic = {}
    with open(os.path.join(OUT, f"{tag}_ncu_top.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` of the tcgen05 kernels\n\n"
                f"source report: gpurun_out/{rep} (scratch, not tracked); selected raw metrics per launch.\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n## {d.get('Kernel Name', '?')}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for i, h in enumerate(hdr):
                if h in KEEP and r[i] not in ("", "n/a"):
                    f.write(f"| {h} | {r[i]} | {units[i]} |\n")
            try:
                rd = float(d["dram__bytes_read.sum"].replace(",", ""))
                wr = float(d["dram__bytes_write.sum"].replace(",", ""))
                ur, uw = units[hdr.index("dram__bytes_read.sum")], units[hdr.index("dram__bytes_write.sum")]
                sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                f.write(f"\nDRAM traffic (read + write) per launch: {(rd * sc[ur] + wr * sc[uw]) / 1e9:.3f} GB\n")
                m = re.search(r"(\w+_kernel)", d.get("Kernel Name", ""))
                if m:
                    traffic[m.group(1)] = {"dram_bytes_per_launch": rd * sc[ur] + wr * sc[uw], "profile": tag,
                                           "ms": d.get("gpu__time_duration.sum")}
            except (KeyError, ValueError):
                pass
    if traffic:       # bench.py reads this for roofline.traffic
        import json
        with open(os.path.join(OUT, "ncu_traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    print("wrote", f"{tag}_ncu_top.md")


if __name__ == "__main__":
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    os.makedirs(OUT, exist_ok=True)
    launches(tag)
    ncu_top(tag)
