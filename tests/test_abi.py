"""The C-ABI shared library: loads, exports every symbol the header declares, and its
host-side argument checking (no GPU, no compute calls)."""
import ctypes
import os
import re

import pytest

import __graft_entry__ as ge
from ciaosr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    ge.build()
    return _lib.load()


def header_functions():
    src = open(os.path.join(ROOT, "include", "ciaosr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ciaosr_[a-z0-9_]+)\s*\(", src)))


def test_exports_match_header(lib):
    names = header_functions()
    assert len(names) >= 10
    assert sorted(_lib.EXPORTS) == names
    for n in names:
        assert hasattr(lib, n), n


def test_abi_version(lib):
    assert lib.ciaosr_abi_version() == _lib.ABI_VERSION
    src = open(os.path.join(ROOT, "include", "ciaosr_b200.h")).read()
    assert f"#define CIAOSR_ABI_VERSION {_lib.ABI_VERSION}" in src
    assert f"#define CIAOSR_MAX_LAYERS {_lib.MAX_LAYERS}" in src


def _desc(c=64, hidden=(256, 256, 256, 256), non_local=True, local_size=2):
    d = _lib.HeadDesc()
    d.abi_version, d.channels, d.feat_unfold = _lib.ABI_VERSION, c, 1
    d.local_size, d.non_local_attn, d.softmax_scale = local_size, int(non_local), 1.0
    dk, dv = 9 * c, 9 * c + (c if non_local else 0)
    fake = 0x1000  # never dereferenced by the host-side calls under test

    def mlp(m, i, o):
        dims = [i] + list(hidden) + [o]
        m.n_layers = len(dims) - 1
        for j, v in enumerate(dims):
            m.dims[j] = v
        for j in range(m.n_layers):
            m.weight[j], m.bias[j] = fake, fake

    mlp(d.imnet_k, dk + 4, dk)
    mlp(d.imnet_v, dv + 4, dv)
    mlp(d.imnet_q, dv, 3)
    if non_local:
        a = d.cs_attn
        a.channels, a.n_scales, a.softmax_scale = c, 1, 10.0
        a.scales[0] = 2
        for f in ("match1_w", "match1_b", "match1_slope", "match2_w", "match2_b", "match2_slope",
                  "assembly_w", "assembly_b", "assembly_slope", "down_w", "down_b", "escape_nan"):
            setattr(a, f, fake)
    return d


def test_plan_and_workspace_sizes(lib):
    d = _desc()
    n = ctypes.c_size_t(0)
    assert lib.ciaosr_plan_bytes(ctypes.byref(d), ctypes.byref(n)) == 0
    # at least the fp32 weights of the three MLPs (1.38 M parameters)
    assert n.value > 1_380_000 * 4
    ws = ctypes.c_size_t(0)
    assert lib.ciaosr_workspace_bytes(ctypes.byref(d), 16, 48, 48, 36864, _lib.ENGINE_SIMT,
                                      ctypes.byref(ws)) == 0
    assert ws.value > 16 * 48 * 48 * 64 * 4


@pytest.mark.parametrize("mutate,needle", [
    (lambda d: setattr(d, "abi_version", 99), "abi_version"),
    (lambda d: setattr(d, "feat_unfold", 0), "feat_unfold"),
    (lambda d: setattr(d, "local_size", 4), "local_size"),
    (lambda d: d.imnet_k.dims.__setitem__(0, 123), "imnet_k"),
    (lambda d: d.cs_attn.scales.__setitem__(0, 3), "multi_scale"),
    (lambda d: setattr(d.imnet_q, "n_layers", 1), "hidden_list"),
])
def test_invalid_descriptors(lib, mutate, needle):
    d = _desc()
    mutate(d)
    n = ctypes.c_size_t(0)
    assert lib.ciaosr_plan_bytes(ctypes.byref(d), ctypes.byref(n)) == _lib.E_INVALID
    assert needle in lib.ciaosr_last_error().decode()


def test_no_cpu_fallback(lib):
    """Without a usable sm_100 device the compute entry points fail loudly."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    d = _desc()
    rc = lib.ciaosr_plan_init(ctypes.byref(d), ctypes.c_void_p(256), 1 << 30, None)
    assert rc == _lib.E_NO_DEVICE
    with pytest.raises(_lib.CiaoSRNativeError):
        _lib.check(rc)


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors in ciaosr_b200/_lib.py against the C compiler's view of include/ciaosr_b200.h
    (the header is plain C: compile a probe with gcc and compare sizes and field offsets)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("gcc not available")
    probes = {
        "ciaosr_mlp_desc": (_lib.MlpDesc, ["n_layers", "dims", "weight", "bias"]),
        "ciaosr_cs_attn_desc": (_lib.CsAttnDesc, ["channels", "n_scales", "scales", "softmax_scale", "match1_w",
                                                  "match2_slope", "assembly_w", "down_w", "down_b", "escape_nan"]),
        "ciaosr_head_desc": (_lib.HeadDesc, ["abi_version", "channels", "feat_unfold", "local_size",
                                             "non_local_attn", "softmax_scale", "imnet_q", "imnet_k", "imnet_v",
                                             "cs_attn"]),
        "ciaosr_rdn_desc": (_lib.RdnDesc, ["abi_version", "mid_channels", "num_layers", "sfe1_w", "sfe2_b",
                                           "dense_w", "lff_b", "gff0_w", "gff1_b"]),
        "ciaosr_linear_desc": (_lib.LinearDesc, ["abi_version", "in_features", "out_features", "weight", "bias"]),
        "ciaosr_conv3x3_desc": (_lib.Conv3x3Desc, ["abi_version", "in_channels", "out_channels", "weight", "bias"]),
    }
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "ciaosr_b200.h"', "int main(void) {"]
    for cname, (_, fields) in probes.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for f in fields:
            lines.append(f'  printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines += ['  printf("CIAOSR_N_STAGES %d\\n", CIAOSR_N_STAGES);', "  return 0;", "}"]
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.check_call([gcc, "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(ln.split() for ln in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, (ctype, fields) in probes.items():
        assert int(got[cname]) == ctypes.sizeof(ctype), cname
        for f in fields:
            assert int(got[f"{cname}.{f}"]) == getattr(ctype, f).offset, (cname, f)
    assert int(got["CIAOSR_N_STAGES"]) == _lib.N_STAGES == len(_lib.STAGE_NAMES)
