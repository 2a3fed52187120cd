"""tcgen05 GEMM + fused epilogue vs a plain torch fp32 reference of the same op (fp16 / bf16 inputs, fp32 accumulate)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SELU_L, SELU_A = 1.0507009873554805, 1.6732632423543772


def ref(a, w, bias, act, residual, gap4):
    y = a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    if act == 1:
        y = y * torch.sigmoid(y)
    elif act == 2:
        y = torch.relu(y)
    elif act == 3:
        y = SELU_L * torch.where(y > 0, y, SELU_A * (torch.exp(y) - 1))
    elif act == 4:
        y = torch.sigmoid(y)
    if residual is not None:
        y = y + residual.float()
    if gap4:
        y = y.view(-1, 4, y.shape[1]).mean(1)
    return y


CASES = [
    # M, N, K, act, residual, out_f32, gap4, block_n
    (1000, 96, 16, 1, False, False, False, 0),        # block2a expand: single k16 step, K < 64 (TMA zero fill)
    (128, 16, 32, 0, False, False, False, 0),         # block1a project
    (520, 24, 144, 0, True, False, False, 0),         # block2b project + residual, N < block_n
    (333, 40, 240, 0, True, False, False, 0),         # ragged M, N=40 -> block_n 48
    (777, 144, 24, 1, False, False, False, 0),        # K=24: second k16 step half zero-filled
    (2048, 1152, 192, 1, False, False, False, 0),     # block6 expand, multiple n tiles
    (1024, 320, 1152, 0, False, False, False, 0),     # block7a project, 18 k-blocks
    (4096, 1280, 320, 1, False, False, True, 0),      # top conv + swish + GAP(2x2)
    (1024, 2048, 1280, 2, False, False, False, 0),    # dense
    (512, 1024, 2048, 3, False, True, False, 0),      # dense_2: selu, fp32 out
    (300, 2048, 2048, 2, False, False, False, 256),   # forced block_n 256 (512 TMEM columns)
    (5, 96, 16, 1, False, False, False, 0),           # tiny M
    (70000, 96, 16, 1, False, False, False, 0),       # many tiles per CTA: exercises ring + TMEM phase wrap
    (1024, 1152, 48, 4, False, False, False, 0),      # SE-expand-like: sigmoid, several n tiles of 64k columns
    (4100, 672, 112, 1, False, False, False, 64),     # forced block_n 64, ragged M, TMA-store clipping
    (999, 112, 672, 0, True, False, False, 48),       # forced narrow tiles (direct-store path with n_tiles > 1)
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("M,N,K,act,res,f32,gap4,bn", CASES)
def test_gemm_matches_torch(kws_lib, M, N, K, act, res, f32, gap4, bn, dtype):
    from multilingual_kws_b200.model import gemm_h16
    g = torch.Generator(device="cuda").manual_seed(M * 7 + N * 3 + K)
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).to(dtype)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(dtype)
    bias = torch.randn(N, device="cuda", generator=g) * 0.2
    residual = (torch.randn(M, N, device="cuda", generator=g)).to(dtype) if res else None
    out = gemm_h16(a, w, bias, act, residual, f32, gap4, bn)
    torch.cuda.synchronize()
    want = ref(a, w, bias, act, residual, gap4)
    assert out.shape == want.shape
    err = (out.float() - want).abs()
    ulp = 2 ** -8 if dtype == torch.bfloat16 else 2 ** -11
    tol = 2e-3 + (0 if f32 else 1) * ulp * want.abs()          # 16-bit output rounding
    assert bool((err <= tol + 1e-3 * want.abs()).all()), f"max err {err.max().item()} at {want.abs().max().item()}"


def test_gemm_no_bias_identity(kws_lib):
    from multilingual_kws_b200.model import gemm_h16
    K = 64
    a = torch.randn(256, K, device="cuda").half()
    w = torch.eye(K, device="cuda").half()
    out = gemm_h16(a, w, None, 0, None, True)
    torch.cuda.synchronize()
    assert torch.equal(out, a.float())


def test_gemm_in_place_residual(kws_lib):
    """The skip connection is applied in place (output buffer == residual buffer), as the embedding runtime does."""
    from multilingual_kws_b200.model import gemm_h16
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 3000, 80, 480
    a = (torch.randn(M, K, device="cuda", generator=g) * 0.5).half()
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).half()
    bias = torch.randn(N, device="cuda", generator=g) * 0.2
    x = torch.randn(M, N, device="cuda", generator=g).half()
    want = ref(a, w, bias, 0, x, False)
    buf = x.clone()
    gemm_h16(a, w, bias, 0, buf, out=buf)
    torch.cuda.synchronize()
    assert bool(((buf.float() - want).abs() <= 2e-3 + 2 ** -10 * want.abs()).all())
