"""Engine-vs-golden diagnostics (GPU box): prints max-abs errors per engine, no asserts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tests.util import build_generator, load_case, max_abs

dev = torch.device("cuda:0")
for name in ["head_c64", "head_c64_nonl0"]:
    meta, a = load_case(name)
    for engine in ["simt", "tcgen05"]:
        g = build_generator(meta, dev, engine=engine)
        if not g.head_plan().engine_supported(engine):
            print(name, engine, "unsupported"); continue
        feat = a["feature"].to(dev)
        if meta["non_local"]:
            nl = g.head_plan().cross_scale_attention(feat, engine=engine)
            print(f"{name} {engine} cs_attn: max-abs {(nl.cpu() - a['nonlocal']).abs().max():.3e} "
                  f"scale {a['nonlocal'].abs().mean():.3f}", flush=True)
        for tag in meta["tags"]:
            coord, cell = a[f"coord_{tag}"].to(dev), a[f"cell_{tag}"].to(dev)
            try:
                pred = g.query_rgb([feat], coord, cell)
                torch.cuda.synchronize()
                ref = a[f"pred_{tag}"]
                err = (pred.cpu() - ref).abs()
                print(f"{name} {engine} {tag}: max-abs {err.max():.3e} mean {err.mean():.3e} ref-scale {ref.abs().mean():.3f} "
                      f"nan={int(torch.isnan(pred).sum())}", flush=True)
            except Exception as exc:
                print(name, engine, tag, "FAILED:", repr(exc)[:300], flush=True)
                raise SystemExit(1)
