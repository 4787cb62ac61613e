"""The fused P2 SwinBlock (blocks_transformer.py:133-171) on tcgen05 / TMEM (csrc/swin_tc.cu) against the mma.sync version of the same
fused kernel (csrc/swin_fused.cu) through the C ABI: identical blobs, identical rounding points -> equal up to accumulation order.
Parity of both against the oracle's layer output is in tests/test_gpu_model.py::test_fused_swin_block_matches_oracle_and_unfused_path."""
import pytest
import torch

import yolo_sod_b200  # noqa: F401

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(1, 14, 14), (2, 30, 23), (3, 56, 56), (5, 160, 160), (1, 8, 9)], ids=lambda s: "x".join(map(str, s)))
@pytest.mark.parametrize("xcs_extra", [0, 64])
def test_swin64_tc_equals_mma_sync(shape, xcs_extra):
    from yolo_sod_b200 import lib
    N, H, W = shape
    g = torch.Generator().manual_seed(7)
    cs = 64 + xcs_extra
    xb = torch.zeros(N, H, W, cs)
    xb[..., xcs_extra:] = torch.randn(N, H, W, 64, generator=g)
    xb = xb.bfloat16().cuda()
    # dw[9][64] | wqkv[192][64] | wo[64][64] | w1[128][64] | w2[64][128] | wpw[64][64]
    wb = torch.cat([torch.randn(576, generator=g) / 3, torch.randn(192 * 64, generator=g) / 8, torch.randn(64 * 64, generator=g) / 8,
                    torch.randn(128 * 64, generator=g) / 8, torch.randn(64 * 128, generator=g) / 11, torch.randn(64 * 64, generator=g) / 8]).bfloat16().cuda()
    pf = torch.randn(768, generator=g) * 0.1
    pf[192:320] = 0.0   # K / V thirds of the in_proj bias: folded away by the caller (ysod.h ysod_swin64_tc), not read by the tcgen05 kernel
    pf = pf.cuda()
    outs = []
    for name in ("ysod_swin64_fused", "ysod_swin64_tc"):
        ob = torch.full((N, H, W, cs), 7.0, dtype=torch.bfloat16, device="cuda")
        lib.call(name, lib.ptr(xb, xcs_extra), N, H, W, cs, lib.ptr(wb), lib.ptr(pf), lib.ptr(ob, xcs_extra), cs, 7, 2, lib.stream_ptr())
        torch.cuda.synchronize()
        if xcs_extra:
            assert bool((ob[..., :xcs_extra] == 7.0).all()), "kernel wrote outside its channel slice"
        outs.append(ob[..., xcs_extra:].float().cpu())
    ref, got = outs
    assert bool(torch.isfinite(got).all())
    d = (got - ref).abs()
    mx = float(ref.abs().max())
    assert float(d.max()) <= 0.03 * mx, (float(d.max()), mx)
    assert float((d > 0.005 * mx).float().mean()) < 2e-3, float((d > 0.005 * mx).float().mean())
    assert float(d.norm() / ref.norm()) < 2e-3, float(d.norm() / ref.norm())
