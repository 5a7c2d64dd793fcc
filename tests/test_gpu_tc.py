"""Pins the tcgen05 building blocks of the tensor-core encoder (UMMA shared-memory / instruction descriptors,
128-byte swizzle, bulk-copied operand image, TMEM read-back) on one 128 x N x K tile product against float64."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dtype,prec", [(torch.bfloat16, 1), (torch.float16, 2)])
@pytest.mark.parametrize("K,N", [(64, 128), (256, 128), (128, 256), (128, 64)])
def test_umma_tile_product(dtype, prec, K, N):
    from jampr_plus_b200 import _cabi
    from jampr_plus_b200.packing import sw128_image
    lib = _cabi.load_library()
    g = torch.Generator().manual_seed(K * 1000 + N)
    A = (torch.rand(128, K, generator=g) * 2 - 1).to(dtype)
    B = (torch.rand(N, K, generator=g) * 2 - 1).to(dtype)
    img = sw128_image(B).cuda()
    Ad = A.cuda()
    D = torch.full((128, N), float("nan"), device="cuda")
    rc = lib.jampr_selftest_umma(_cabi.ptr(Ad), _cabi.ptr(img), _cabi.ptr(D), K, N, prec, _cabi.stream_ptr())
    assert rc == 0
    torch.cuda.synchronize()
    ref = A.double() @ B.double().t()
    err = (D.cpu().double() - ref).abs().max().item()
    assert err < 1e-4 * K, err      # exact products of 16-bit operands, fp32 accumulation


def test_sw128_image_layout():
    from jampr_plus_b200.packing import sw128_image
    w = torch.arange(16 * 128, dtype=torch.int16).view(16, 128).view(torch.bfloat16)
    img = sw128_image(w).view(torch.int16).view(2, 16, 8, 8)
    src = w.view(torch.int16).view(16, 2, 8, 8)
    for kb in range(2):
        for r in range(16):
            for c in range(8):
                assert torch.equal(img[kb, r, c ^ (r & 7)], src[r, kb, c])
