"""-m gpu: single-layer checks of the non-GEMM forward kernels (csrc/model_kernels.cu) against
plain PyTorch fp32 references of the same op, and of the fused variants against their unfused
compositions (bit-exact: the fusions keep the same rounding points)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    from empanada_napari_b200 import pdl  # noqa: F401  (declares the be_op_* signatures)
    from empanada_napari_b200._lib import call, ptr
    return torch, call, ptr


@pytest.mark.parametrize("B,H,W,C,k", [(2, 64, 64, 256, 5), (1, 37, 51, 72, 5), (3, 16, 24, 128, 3), (1, 9, 70, 8, 3)])
def test_dwconv_vs_torch(B, H, W, C, k):
    torch, call, ptr = _setup()
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(B * 1000 + C + k)
    x = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(dev)
    w = torch.randn(C, 1, k, k, generator=g) * 0.2
    wt = w.reshape(C, k * k).t().contiguous().to(dev)          # [k*k][C] fp32
    out = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=dev)
    call("be_op_dwconv", None, ptr(x), C, B, H, W, C, k, ptr(wt), ptr(out), C, None, 0, 0, 0, None)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.to(dev), padding=k // 2, groups=C)
    ref = ref.permute(0, 2, 3, 1)
    err = (out.float() - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 1e-3      # one bf16 rounding of an fp32-accumulated sum
    assert bool((err <= tol).all()), float(err.max())


@pytest.mark.parametrize("B,Hu,Wu,clow", [(2, 8, 8, 32), (1, 5, 7, 16)])
def test_dwconv_fused_upsample_equals_unfused(B, Hu, Wu, clow):
    """bilinear(align_corners=True) + concat + depthwise 5x5 in one kernel == the three steps."""
    torch, call, ptr = _setup()
    dev = torch.device("cuda:0")
    H, W, Cup = 4 * Hu, 4 * Wu, 256
    C = Cup + clow
    g = torch.Generator(device="cpu").manual_seed(7)
    up = torch.randn(B, Hu, Wu, Cup, generator=g).to(torch.bfloat16).to(dev)
    low = torch.randn(B, H, W, clow, generator=g).to(torch.bfloat16).to(dev)
    wt = (torch.randn(25, C, generator=g) * 0.2).to(dev)
    fused = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=dev)
    call("be_op_dwconv", None, ptr(low), clow, B, H, W, C, 5, ptr(wt), ptr(fused), C, ptr(up), Cup, Hu, Wu, None)
    cat = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=dev)
    call("be_op_bilinear", None, ptr(up), Cup, B, Hu, Wu, Cup, ptr(cat), C, 0, H, W, None)
    cat[..., Cup:] = low
    plain = torch.zeros_like(fused)
    call("be_op_dwconv", None, ptr(cat), C, B, H, W, C, 5, ptr(wt), ptr(plain), C, None, 0, 0, 0, None)
    torch.cuda.synchronize()
    assert torch.equal(fused, plain)
    # and the bilinear producer itself against torch
    ref = torch.nn.functional.interpolate(up.float().permute(0, 3, 1, 2), size=(H, W), mode="bilinear", align_corners=True)
    got = cat[..., :Cup].float().permute(0, 3, 1, 2)
    assert float((got - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max()) + 1e-3


@pytest.mark.parametrize("B,h,w,axis,dtype", [(2, 64, 64, 0, "uint8"), (1, 100, 72, 0, "uint8"), (2, 48, 80, 1, "uint8"),
                                              (2, 80, 48, 2, "uint8"), (2, 48, 80, 1, "uint16"), (1, 64, 48, 2, "int16"),
                                              (1, 48, 64, 0, "int32"), (1, 48, 64, 0, "int8")])
def test_stem_pool_equals_stem_then_maxpool(B, h, w, axis, dtype):
    """Fused conv1+BN+ReLU+MaxPool kernel == stem kernel followed by the max-pool kernel (bits),
    and both agree with torch conv2d + max_pool2d on the normalised, padded slice."""
    torch, call, ptr = _setup()
    dev = torch.device("cuda:0")
    pf = 16
    H, W = h + (pf - h % pf) % pf, w + (pf - w % pf) % pf
    g = torch.Generator(device="cpu").manual_seed(11)
    shape = [B + 1, h, w]
    shape3d = {0: (B + 1, h, w), 1: (h, B + 1, w), 2: (h, w, B + 1)}[axis]
    from empanada_napari_b200.pdl import elem_code, norm_constants
    info = np.iinfo(np.dtype(dtype))
    # values spread over the dtype's whole range (the reference normalises by iinfo.max, utils.py:189-201)
    vol_np = (torch.rand(shape3d, generator=g, dtype=torch.float64).numpy() * (float(info.max) - float(info.min)) + float(info.min)).astype(dtype)
    vol = torch.from_numpy(vol_np).to(dev)
    elem = elem_code(vol.dtype)
    D, Hv, Wv = shape3d
    strides = [(Hv * Wv, Wv, 1), (Wv, Hv * Wv, 1), (1, Hv * Wv, Wv)][axis]
    wt = (torch.randn(64, 1, 7, 7, generator=g) * 0.1)
    bias = torch.randn(64, generator=g) * 0.1
    wt_d = wt.reshape(64, 49).t().contiguous().to(dev)
    bias_d = bias.to(dev)
    mean255, den = (float(v) for v in norm_constants({"mean": 0.57571, "std": 0.12765}, vol.dtype))
    s0 = 1
    stem = torch.zeros(B, H // 2, W // 2, 64, dtype=torch.bfloat16, device=dev)
    pooled = torch.zeros(B, H // 4, W // 4, 64, dtype=torch.bfloat16, device=dev)
    fused = torch.zeros_like(pooled)
    call("be_op_stem", None, B, h, w, H, W, mean255, den, ptr(wt_d), ptr(bias_d), ptr(stem), ptr(vol), *strides, s0, elem, None)
    call("be_op_maxpool", None, ptr(stem), B, H // 2, W // 2, 64, ptr(pooled), H // 4, W // 4, None)
    call("be_op_stem_pool", None, B, h, w, H, W, mean255, den, ptr(wt_d), ptr(bias_d), ptr(fused), ptr(vol), *strides, s0, elem, None)
    torch.cuda.synchronize()
    assert torch.equal(fused, pooled)
    sl = torch.from_numpy(np.moveaxis(vol_np, axis, 0)[s0:s0 + B].astype(np.float32)).to(dev)
    x = torch.zeros(B, 1, H, W, device=dev)
    x[:, 0, :h, :w] = (sl - mean255) * den
    ref = torch.nn.functional.conv2d(x, wt.to(dev), bias_d, stride=2, padding=3).relu()
    ref = torch.nn.functional.max_pool2d(ref, 3, 2, 1).permute(0, 2, 3, 1)
    err = (fused.float() - ref).abs()
    assert float(err.max()) <= 2.0 ** -7 * float(ref.abs().max()) + 2e-3, float(err.max())


@pytest.mark.parametrize("B,Hi,Wi,Cin", [(2, 1, 2, 128), (1, 8, 16, 256), (3, 5, 3, 256)])
def test_convt2x2_vs_torch(B, Hi, Wi, Cin):
    """ConvTranspose2d(k=2, s=2) + bias + ReLU as one GEMM with a pixel-shuffle epilogue, written
    into the first half of a 2*Cout-channel concat buffer (decoders/bifpn.py:226-234)."""
    torch, call, ptr = _setup()
    dev = torch.device("cuda:0")
    Cout = 128
    g = torch.Generator(device="cpu").manual_seed(Cin + Hi)
    x = torch.randn(B, Hi, Wi, Cin, generator=g).to(torch.bfloat16).to(dev)
    w = (torch.randn(Cin, Cout, 2, 2, generator=g) / Cin ** 0.5).to(torch.bfloat16)
    bias = torch.randn(Cout, generator=g) * 0.1
    wg = w.permute(2, 3, 1, 0).reshape(4 * Cout, Cin).contiguous().to(dev)
    bias4 = bias.repeat(4).to(dev)
    out = torch.full((B, 2 * Hi, 2 * Wi, 2 * Cout), 7.0, dtype=torch.bfloat16, device=dev)
    call("be_op_convt2x2", None, ptr(x), Cin, B, Hi, Wi, Cin, ptr(wg), Cout, ptr(out), 2 * Cout, 0, ptr(bias4), 1, None)
    torch.cuda.synchronize()
    ref = torch.nn.functional.conv_transpose2d(x.float().permute(0, 3, 1, 2), w.float().to(dev), bias.to(dev), stride=2).relu()
    ref = ref.permute(0, 2, 3, 1)
    got = out[..., :Cout].float()
    assert float((got - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max()) + 2e-3
    assert bool((out[..., Cout:] == 7.0).all())      # the skip half of the concat buffer is untouched


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_bifpn_fuse_vs_torch(mode):
    torch, call, ptr = _setup()
    dev = torch.device("cuda:0")
    B, H, W, C = 2, 6, 10, 128
    Ha, Wa = {0: (H, W), 1: (H // 2, W // 2), 2: (2 * H, 2 * W - 1)}[mode]
    g = torch.Generator(device="cpu").manual_seed(mode)
    a = torch.randn(B, Ha, Wa, C, generator=g).to(torch.bfloat16).to(dev)
    bcat = torch.randn(B, H, W, 2 * C, generator=g).to(torch.bfloat16).to(dev)   # b lives in a concat slot
    c = torch.randn(B, H, W, C, generator=g).to(torch.bfloat16).to(dev)
    w1, w2, w3 = 0.31, 0.42, 0.27
    denom = w1 + w2 + w3 + 1e-4
    out = torch.zeros(B, H, W, C, dtype=torch.bfloat16, device=dev)
    call("be_op_bifpn_fuse", None, ptr(a), C, mode, Ha, Wa, ptr(bcat[..., C:]), 2 * C, ptr(c), C,
         w1, w2, w3, denom, B, H, W, C, ptr(out), C, None)
    torch.cuda.synchronize()
    an = a.float().permute(0, 3, 1, 2)
    if mode == 1:
        an = torch.nn.functional.interpolate(an, scale_factor=2.0, mode="nearest")
    elif mode == 2:
        an = torch.nn.functional.max_pool2d(an, 3, 2, 1)
    ref = (w1 * an.permute(0, 2, 3, 1) + w2 * bcat[..., C:].float() + w3 * c.float()) / denom
    assert float((out.float() - ref).abs().max()) <= 2.0 ** -7 * float(ref.abs().max()) + 1e-3
