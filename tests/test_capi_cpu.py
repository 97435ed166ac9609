"""The C-ABI library loads without a GPU and exports every symbol include/nunet_b200.h declares; the host-only plan
functions accept every layer shape of the BASELINE configs; bad descriptors are refused with error codes."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nunet_b200.h")).read()
    return sorted(set(re.findall(r"\b(?:int|long long)\s+(nunet_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import _lib as L
    lib = L.lib()
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    assert set(L.exported_symbols()) == set(names)
    assert lib.nunet_version() == L.ABI_VERSION


def _layers(size, N):
    import engine
    nb = engine.NB_FILTER
    out = []
    for graph in (engine.nested_graph(), engine.unet_graph()):
        for node, srcs in graph:
            lvl = int(node[0])
            cin = sum(16 if k == "input" else nb[int(s[0])] for s, k in srcs)
            out.append((N, size >> lvl, size >> lvl, cin, nb[lvl]))
            out.append((N, size >> lvl, size >> lvl, nb[lvl], nb[lvl]))
    return sorted(set(out))


@pytest.mark.parametrize("size,N", [(96, 256), (96, 16), (256, 64), (512, 8), (96, 1), (96, 1024), (32, 2)])
def test_plans_cover_all_baseline_layer_shapes(size, N):
    import _lib as L
    lib = L.lib()
    for (n, h, w, cin, cout) in _layers(size, N):
        pl = L.ConvPlan()
        assert lib.nunet_conv3x3_plan(n, h, w, cin, cout, C.byref(pl)) == 0, (n, h, w, cin, cout)
        assert pl.smem_bytes <= 227 * 1024 and pl.T * pl.NT * 2 <= 512 and pl.tmem_cols <= 512
        assert pl.NT * pl.n_tiles == cout and pl.nchunks * pl.CK == pl.cin_pad and 1 <= pl.S <= 3  # the stage ring spans tile groups
        # rows wider than a tensor-load box are planned as column strips (tensor-load sources, 32-channel chunks)
        strips = max(pl.strips, 1)
        ws = (w + strips - 1) // strips
        ws += (ws & 1) if strips > 1 else 0
        if strips > 1:
            assert w + 2 > 256 and cin % 32 == 0 and ws + 2 <= 256
            assert pl.CK == 32 and 0 < pl.tma_smem_bytes <= 227 * 1024
        assert pl.P == 128 * pl.T + 2 * (ws + 2) + 2 and pl.P * 16 < (1 << 18)
        assert pl.grid_m * pl.n_tiles <= 148 and pl.grid_m <= pl.ntg
        if cin >= 32:  # dgrad runs the same kernel with the roles of cin / cout swapped
            pd = L.ConvPlan()
            assert lib.nunet_conv3x3_plan(n, h, w, cout, cin, C.byref(pd)) == 0, ("dgrad", n, h, w, cout, cin)
            assert pd.smem_bytes <= 227 * 1024 and pd.T * pd.NT * 2 <= 512
        pw = L.WgradPlan()
        assert lib.nunet_conv3x3_wgrad_plan(n, h, w, cin, cout, C.byref(pw)) == 0
        assert pw.smem_bytes <= 227 * 1024 and pw.KP % 16 == 0 and pw.splits * pw.cpc >= pw.nchunks_total
        assert (pw.splits - 1) * pw.cpc < pw.nchunks_total  # no empty split


def test_bad_arguments_are_refused_without_cuda():
    import _lib as L
    lib = L.lib()
    pl = L.ConvPlan()
    assert lib.nunet_conv3x3_plan(0, 96, 96, 32, 32, C.byref(pl)) == 1          # NUNET_E_BADARG
    assert lib.nunet_conv3x3_plan(4, 96, 96, 32, 40, C.byref(pl)) == 2          # Cout % 32 != 0 -> unsupported
    assert lib.nunet_conv3x3_plan(4, 96, 96, 24, 32, C.byref(pl)) == 2          # cin not a multiple of 16
    assert lib.nunet_conv3x3_fwd(None, None) == 1
    assert lib.nunet_conv3x3_wgrad(None, None) == 1
    assert lib.nunet_grad_router(None, None) == 1
    d = L.ConvDesc()
    d.num_src = 9
    assert lib.nunet_conv3x3_fwd(C.byref(d), None) == 1
    assert lib.nunet_head_fwd(None, None, None, None, None, 1, 8, 8, 1, None, None) == 1
    assert lib.nunet_loss_iou_fwd(None, None, 1, 1, None, None, None, None, None, None) == 1


@pytest.mark.parametrize("size,N", [(96, 256), (96, 16), (256, 64), (512, 8), (32, 2)])
def test_tma_weight_gradient_and_stacked_conv_tilings(size, N):
    """Host-only tiling decisions of the TMA-fed kernels: every eligible layer shape gets a tiling that fits shared memory,
    covers all chunks, uses the kernel-row stacking only for Cin = 32, and rows wider than 254 pixels are left to the cp.async
    kernels; the stacked Cout = 32 conv kernel is offered exactly for 32 output / >= 64 input channels."""
    import _lib as L
    lib = L.lib()
    for (n, h, w, cin, cout) in _layers(size, N):
        pw = L.WgradPlan()
        assert lib.nunet_conv3x3_wgrad_plan(n, h, w, cin, cout, C.byref(pw)) == 0
        want_tma = cin % 32 == 0          # rows wider than 254 px: column strips (rows mode only)
        assert bool(pw.tma) == want_tma, (n, h, w, cin, cout)
        strips = (w + 253) // 254 if w + 2 > 256 else 1
        ws = (w + strips - 1) // strips
        ws += (ws & 1) if strips > 1 else 0
        nv, wv = n * strips, ws               # the tiling below is that of the strip geometry
        if pw.tma:
            assert pw.tma_smem_bytes <= 227 * 1024 and 1 <= pw.tma_S <= 4
            assert pw.tma_PW >= wv + 2 and pw.tma_PW % 2 == 0 and pw.tma_PW <= 256 and (strips == 1 or pw.tma_r > 0)
            assert pw.tma_splits * pw.tma_cpc >= pw.tma_chunks and (pw.tma_splits - 1) * pw.tma_cpc < pw.tma_chunks
            assert pw.tma_G == (1 if cin == 32 else 3) and pw.tma_Nt == 32
            assert pw.part_elems >= pw.tma_splits * 9 * cin * cout
            # chunk enumeration covers every padded row exactly once: ry rows of one image, or kimg whole images
            if pw.tma_r > 0:
                assert pw.tma_chunks == nv * ((h + 1 + pw.tma_r - 1) // pw.tma_r)
                assert -(-(pw.tma_r * pw.tma_PW) // 16) * 16 <= 1024
            else:
                assert pw.tma_chunks == (nv + (-pw.tma_r) - 1) // (-pw.tma_r) and (-pw.tma_r) * (h + 1) <= 256
            m_tiles = (cin + 127) // 128
            assert pw.tma_splits * m_tiles * (cout // 32) <= 148 or pw.tma_splits == 1     # one wave of CTAs
        want_c32 = cout == 32 and cin >= 64 and cin % 32 == 0      # any row width (column strips above 254 px)
        assert bool(lib.nunet_conv3x3_c32_eligible(n, h, w, cin, cout)) == want_c32, (n, h, w, cin, cout)
