"""gpurun_out/configs*.jsonl (tools/run_configs.py) -> profiles/<tag>_configs.md."""
import glob
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
    files = sys.argv[2:] or sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "configs*.jsonl")))
    rows = []
    for f in files:
        for ln in open(f):
            if ln.startswith("{"):
                rows.append(json.loads(ln))
    out = os.path.join(ROOT, "profiles", f"{tag}_configs.md")
    with open(out, "w") as f:
        f.write(f"# {tag}: BASELINE.json configs 3-5 at full size (tools/run_configs.py; synthetic weights and images)\n\n"
                "Timed call = the restorer's inference path (encoder + head + tile blend + de-normalise/clamp), result on "
                "the device, CUDA events, best of 2 after 1 warm-up; generator forward replayed as a CUDA graph per tile "
                "shape.  `check` = product path (tcgen05 engine, native RDN encoder) vs the all-fp32 CUDA-core engine + "
                "PyTorch fp32 encoder on a crop of the same input (max-abs / PSNR between the two outputs; tolerance "
                "1e-4).  Stage times are one eager pass of the whole frame.\n\n"
                "| config | case | GPUs | tiles | HR px | ms | Mpix/s | peak GB | check max-abs | check PSNR dB | stages (ms per frame) |\n"
                "|---|---|---:|---:|---:|---:|---:|---:|---:|---:|---|\n")
        for r in rows:
            st = r.get("stage_ms_per_frame")
            st = ", ".join(f"{k} {v:.1f}" for k, v in st.items()) if st else ""
            ck = f"{r['check_max_abs']:.1e}" if "check_max_abs" in r else ""
            ps = f"{r['check_psnr_db']:.1f}" if "check_psnr_db" in r else ""
            f.write(f"| {r['config']} | {r['case']} | {r['n_gpus']} | {r['tiles']} | {r['hr_px']} | {r['ms']:.1f} | "
                    f"{r['mpix_s']:.2f} | {r['peak_mem_gb']:.1f} | {ck} | {ps} | {st} |\n")
    print("wrote", out)


if __name__ == "__main__":
    main()
