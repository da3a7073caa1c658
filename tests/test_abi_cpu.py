"""CPU-side checks: the C-ABI library loads and exports every symbol include/inrf.h declares,
argument validation works without a GPU, the Python layer refuses CPU tensors."""
import ctypes as C
import os

import pytest
import torch

import intrinsicnerf_b200 as inrf
from intrinsicnerf_b200 import _lib


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    declared = _lib.declared_symbols()
    assert len(declared) >= 20
    missing = [s for s in declared if not hasattr(L, s)]
    assert not missing, missing
    assert set(_lib._SIGS) == set(declared)
    assert L.inrf_version() >= 100


def test_layout_queries_and_rejections():
    L = _lib.lib()
    # 662152 parameters per object network, 698660 for SSR with C=28 (M2)
    assert L.inrf_flat_param_count(_lib.NET_OBJECT, 0) == 662152
    assert L.inrf_flat_param_count(_lib.NET_SSR, 28) == 698660
    assert L.inrf_packed_bytes(_lib.NET_OBJECT, 0) > 662152 * 4
    assert L.inrf_flat_param_count(_lib.NET_OBJECT, 5) < 0          # object net has no semantic head
    assert L.inrf_flat_param_count(7, 0) < 0
    assert b"variant" in L.inrf_last_error_string()
    assert L.inrf_flat_param_count(_lib.NET_SSR, 500) == -2          # INRF_EUNSUPPORTED


def test_null_pointers_are_rejected_not_dereferenced():
    L = _lib.lib()
    assert L.inrf_embed(None, 5, 10, 1.0, None, None) == -1
    assert L.inrf_merge_sorted(None, None, 4, 64, 128, None, None, None) == -1
    cfg = _lib.RenderCfg(variant=0, n_classes=0, n_samples=64, n_importance=128, pe_scalar_factor=1.0)
    assert L.inrf_render_workspace_bytes(C.byref(cfg), 1024) > 1024 * 192 * 11 * 4
    cfg.n_samples = 4000
    assert L.inrf_render_workspace_bytes(C.byref(cfg), 1024) == -2
    cfg = _lib.RenderCfg(variant=1, n_classes=28, n_samples=64, n_importance=128, lindisp=1, pe_scalar_factor=10.0)
    assert L.inrf_render_workspace_bytes(C.byref(cfg), 8) == -2      # SSR has no lindisp
    # zero-size calls are no-ops even with null pointers
    assert L.inrf_embed(None, 0, 10, 1.0, None, None) == 0


def test_no_cpu_fallback():
    with pytest.raises(RuntimeError, match="CUDA"):
        inrf.ops.embed(torch.zeros(4, 3), 10)
    net = inrf.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True)
    with pytest.raises(RuntimeError, match="CUDA"):
        with torch.no_grad():
            net(torch.zeros(2, 90))
    with pytest.raises(NotImplementedError):
        inrf.NeRF(D=4, W=128, input_ch=63, input_ch_views=27, use_viewdirs=True)
    with pytest.raises(NotImplementedError):
        inrf.get_embedder(10, -1)


def test_state_dict_layout_matches_reference_names():
    torch.manual_seed(0)
    net = inrf.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True)
    keys = list(net.state_dict().keys())
    assert keys[:2] == ["pts_linears.0.weight", "pts_linears.0.bias"]
    assert keys[16:18] == ["views_linears.0.weight", "views_linears.0.bias"]
    assert keys[18:] == [f"{n}.{p}" for n in ("feature_linear", "alpha_linear", "shading_linear", "albedo_linear1",
                                              "albedo_linear2", "test_linear1", "test_linear2") for p in ("weight", "bias")]
    assert sum(p.numel() for p in net.parameters()) == 662152
    assert net.flat_params().numel() == 662152
    ssr = inrf.Semantic_NeRF(True, 28, D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True)
    assert "semantic_linear.0.0.weight" in ssr.state_dict() and "semantic_linear.1.bias" in ssr.state_dict()
    assert sum(p.numel() for p in ssr.parameters()) == 698660 == ssr.flat_params().numel()


def test_seeded_init_equals_oracle_init():
    """Same construction order as the reference => same weights from the same seed."""
    from oracle import nerf_oracle as orc
    torch.manual_seed(20220414)
    net = inrf.NeRF(D=8, W=256, input_ch=63, input_ch_views=27, use_viewdirs=True)
    torch.manual_seed(20220414)
    p = orc.init_params("object")
    for k, v in net.state_dict().items():
        assert torch.equal(v, p[k]), k


def test_struct_layouts_match_header(tmp_path):
    """The ctypes mirrors of InrfRenderCfg / InrfFramePlanes have the size and field offsets a C compiler gives the
    header's structs (include/inrf.h is plain C: it must also compile as C, not only as C++/CUDA)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    fields_cfg = [n for n, _ in _lib.RenderCfg._fields_]
    fields_pl = [n for n, _ in _lib.FramePlanes._fields_]
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{_lib.HEADER}"', 'int main(void) {',
             '  printf("%zu %zu\\n", sizeof(InrfRenderCfg), sizeof(InrfFramePlanes));']
    lines += [f'  printf("%zu\\n", offsetof(InrfRenderCfg, {f}));' for f in fields_cfg]
    lines += [f'  printf("%zu\\n", offsetof(InrfFramePlanes, {f}));' for f in fields_pl]
    lines += ['  return 0;', '}']
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-o", str(exe), str(src)])
    out = subprocess.check_output([str(exe)], text=True).split()
    assert [int(out[0]), int(out[1])] == [C.sizeof(_lib.RenderCfg), C.sizeof(_lib.FramePlanes)]
    want = [getattr(_lib.RenderCfg, f).offset for f in fields_cfg] + [getattr(_lib.FramePlanes, f).offset for f in fields_pl]
    assert [int(x) for x in out[2:]] == want


def test_split_rec_gradient_equals_slicing():
    """ops.SplitRecFn (one concatenation in the backward) == autograd through plain column slices of the record."""
    import torch
    from intrinsicnerf_b200 import ops
    torch.manual_seed(0)
    for C, ep in ((0, False), (5, False), (5, True)):
        W = 13 + C + (128 if ep else 0)
        rec = torch.randn(9, W, requires_grad=True)
        m = ops.split_rec(rec, C, ep)
        loss = (m["rgb"] ** 2).sum() + m["shading"].sum() * 3 + (m["albedo"] * 2).sum() + (m["sem"].sum() if C else 0) + (m["feat"].mean() if ep else 0)
        loss.backward()
        ref = rec.detach().clone().requires_grad_(True)
        want = (ref[:, 0:3] ** 2).sum() + ref[:, 8].sum() * 3 + (ref[:, 5:8] * 2).sum() + (ref[:, 13:13 + C].sum() if C else 0) + (ref[:, 13 + C:].mean() if ep else 0)
        want.backward()
        assert torch.equal(rec.grad, ref.grad)
        assert torch.equal(m["disp"], rec.detach()[:, 3]) and m["rgb"].shape == (9, 3)
        with torch.no_grad():
            assert torch.equal(ops.split_rec(rec, C, ep)["residual"], rec[:, 9:12])
