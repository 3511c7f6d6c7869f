"""CPU-side checks: the C-ABI library loads, exports every symbol include/cope_b200.h declares, the product path
refuses to run without CUDA (no fallback), and host-side logic (constructors, state_dict names, flat layout)."""
import os
import re
import subprocess

import pytest
import torch

import cope_nerf_b200 as C
from cope_nerf_b200 import _lib as L
from conftest import ROOT, load_golden


def _declared():
    src = open(os.path.join(ROOT, "include", "cope_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cope_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = C.load_library()
    names = _declared()
    assert len(names) >= 28
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (cope_[a-z0-9_]+)", out))
    for n in names:
        assert n in exported, f"{n} declared in include/cope_b200.h but not exported"
        assert hasattr(lib, n)
    assert set(L.EXPORTS) == set(names), set(L.EXPORTS) ^ set(names)
    assert lib.cope_version() >= 100


def test_library_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "-lelf", L.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback():
    torch.manual_seed(0)
    r = C.training.build_networks(device="cpu")
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises((C.CopeError, RuntimeError)):
        r.sdf_network(torch.zeros(4, 4))
    with pytest.raises((C.CopeError, RuntimeError)):
        C.make_c2w(torch.zeros(3), torch.zeros(3))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "cope_nerf_b200")
    for f in os.listdir(pkg):
        if f.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, f)).read().replace("# oracle", ""), f


def test_constructors_match_reference_rng_stream_and_keys():
    g = load_golden("init_seed678")
    torch.manual_seed(678)
    cfg = C.training.DEFAULT_CFG
    sdf = C.SDFNetwork(**cfg["neus_sdf_network"])
    col = C.RenderingNetwork(**cfg["neus_rendering_network"])
    var = C.SingleVarianceNetwork(**cfg["neus_variance_network"])
    for tag, m in (("sdf", sdf), ("color", col), ("variance", var)):
        sd = m.state_dict()
        want = {k[len(tag) + 1:-4] for k in g if k.startswith(tag + ".") and k.endswith(".sum")}
        assert set(sd.keys()) == want
        for k, v in sd.items():
            assert torch.equal(v.flatten()[:4], g[f"{tag}.{k}.head"].flatten()), (tag, k)
            assert abs(v.double().sum().item() - g[f"{tag}.{k}.sum"].item()) < 1e-9
    assert sdf.n_flat == L.query("cope_mlp_flat_floats", sdf.desc) == sum(
        a * b + b for a, b in zip(sdf._dims_in, sdf._dims_out))
    assert sdf._dims_in == [52, 256, 256, 256, 256, 256, 256, 256, 256]
    assert sdf._dims_out == [256, 256, 256, 204, 256, 256, 256, 256, 257]
    assert col._dims_in == [291, 256, 256, 256, 256] and col._dims_out == [256, 256, 256, 256, 3]


def test_renderer_state_dict_names_are_checkpoint_compatible():
    r = C.training.build_networks(device="cpu")
    keys = set(r.state_dict().keys())
    assert "sdf_network.lin0.weight_g" in keys and "color_network.lin4.weight_v" in keys
    assert "deviation_network.variance" in keys
    p = C.PoseRetriever(3)
    assert set(p.state_dict().keys()) == {"init_c2w", "r", "t"}


def test_patch_indices_match_reference():
    g = load_golden("poses_rays")
    torch.manual_seed(9)
    idx = C.training.get_patch_indices(int(g["H"]), int(g["W"]), 4, 64)
    assert torch.equal(idx, g["idx"])
    pix = C.pixels_from_indices(idx, int(g["H"]), int(g["W"]))
    assert torch.equal(pix, g["pix"])
