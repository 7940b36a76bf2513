"""GPU parity of the tcgen05 GEMM family (csrc/gemm.cu) against fp32 torch matmuls on the same bf16-rounded
operands.  Tolerances: fp32 outputs rel 2e-3 of the output scale (bf16 products, fp32 accumulate; summation order
differs); bf16 outputs additionally carry one bf16 rounding (2^-8 relative)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(rows, cols, pad=8, scale=1.0, seed=0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ld = (cols + pad - 1) // pad * pad
    buf = torch.randn(rows, ld, generator=g, device="cuda") * scale
    buf = buf.to(dtype)
    return buf[:, :cols]


def _close(out, ref, tol, what):
    out = out.float()
    ref = ref.float()
    scale = ref.abs().max().item() + 1e-6
    err = (out - ref).abs().max().item()
    assert err <= tol * scale, f"{what}: max err {err:.4g} vs scale {scale:.4g} (tol {tol})"


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 64, 64), (300, 256, 256), (1600, 1336, 668),
                                   (3200, 768, 256), (200, 40, 512), (77, 24, 40)])
@pytest.mark.parametrize("out_fp32", [True, False])
def test_gemm_tn_bias(M, N, K, out_fp32):
    from multi_modal_foundation_model_b200 import ops
    A, B = _mk(M, K, seed=1), _mk(N, K, seed=2)
    bias = torch.randn(N, device="cuda")
    D = torch.full((M, (N + 7) // 8 * 8), 7.0, device="cuda", dtype=torch.float32 if out_fp32 else torch.bfloat16)[:, :N]
    ops.gemm_tn(A, B, D, M=M, N=N, K=K, bias=bias)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().T + bias
    _close(D, ref, 2e-3 if out_fp32 else 8e-3, f"gemm_tn {M}x{N}x{K}")


def test_gemm_tn_epilogues():
    from multi_modal_foundation_model_b200 import ops
    from multi_modal_foundation_model_b200._lib import (ACT_DGELU, ACT_DSOFTSIGN, ACT_GELU, ACT_GELU_DG, ACT_MULAUX,
                                                        ACT_SOFTSIGN)
    M, N, K = 400, 512, 256
    A, B = _mk(M, K, seed=3, scale=0.5), _mk(N, K, seed=4, scale=0.1)
    bias = torch.randn(N, device="cuda") * 0.1
    v = A.float() @ B.float().T + bias
    # GELU + saved pre-activation
    D = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    D2 = torch.empty_like(D)
    ops.gemm_tn(A, B, D, bias=bias, act=ACT_GELU, D2=D2)
    _close(D, torch.nn.functional.gelu(v), 8e-3, "gelu")
    _close(D2, v, 8e-3, "gelu pre-activation")
    # GELU + saved derivative (what the engine records), and its backward epilogue D = v * aux
    ops.gemm_tn(A, B, D, bias=bias, act=ACT_GELU_DG, D2=D2)
    vg = v.clone().requires_grad_(True)
    torch.nn.functional.gelu(vg).sum().backward()
    _close(D, torch.nn.functional.gelu(v), 8e-3, "gelu (dg flavour)")
    _close(D2, vg.grad, 8e-3, "gelu derivative")
    auxm = _mk(M, N, seed=8)
    Dm = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_tn(A, B, Dm, act=ACT_MULAUX, aux=auxm)
    _close(Dm, (A.float() @ B.float().T) * auxm.float(), 8e-3, "mulaux")
    # softsign * scale
    ops.gemm_tn(A, B, D, bias=bias, act=ACT_SOFTSIGN, act_scale=1.5)
    _close(D, torch.nn.functional.softsign(v) * 1.5, 8e-3, "softsign")
    # dgelu: D = v * gelu'(aux)
    aux = _mk(M, N, seed=5)
    xa = aux.float().requires_grad_(True)
    torch.nn.functional.gelu(xa).sum().backward()
    Df = torch.empty(M, N, device="cuda")
    ops.gemm_tn(A, B, Df, act=ACT_DGELU, aux=aux)
    _close(Df, (A.float() @ B.float().T) * xa.grad, 3e-3, "dgelu")
    # dsoftsign through the saved output a = s * x/(1+|x|)
    s = 1.5
    xs = _mk(M, N, seed=6).float()
    a_saved = (torch.nn.functional.softsign(xs) * s).to(torch.bfloat16)
    t = 1.0 - (a_saved.float() / s).abs()
    ops.gemm_tn(A, B, Df, act=ACT_DSOFTSIGN, aux=a_saved, act_scale=s)
    _close(Df, (A.float() @ B.float().T) * s * t * t, 3e-3, "dsoftsign")
    # residual + remap + token zeroing: rows (b, t) -> (b, off + t) of a (B, S, N) buffer
    Bb, T, S, off = 4, 100, 200, 100
    zero = torch.zeros(S, dtype=torch.uint8, device="cuda")
    zero[off + 3] = 1
    zero[off + 50] = 1
    res = torch.randn(Bb * S, N, device="cuda")
    out = torch.zeros(Bb * S, N, device="cuda")
    ops.gemm_tn(A, B, out, bias=bias, res=res, remap=(T, S, off), row_zero=zero)
    ref = torch.zeros(Bb, S, N, device="cuda")
    tok = v.view(Bb, T, N).clone()
    tok[:, 3] = 0
    tok[:, 50] = 0
    ref[:, off:off + T] = tok + res.view(Bb, S, N)[:, off:off + T]
    _close(out, ref.view(Bb * S, N), 3e-3, "remap/zero/res")
    assert out.view(Bb, S, N)[:, :off].abs().max().item() == 0.0


def test_gemm_tn_dropout_matches_stream():
    from multi_modal_foundation_model_b200 import ops
    from oracle import philox_ref as px
    M, N, K = 256, 200, 64
    A, B = _mk(M, K, seed=7), _mk(N, K, seed=8)
    seed = torch.tensor([0x1234567890ABCDEF], dtype=torch.int64, device="cuda")
    D = torch.empty(M, N, device="cuda")
    ops.gemm_tn(A, B, D, drop=ops.DropSpec(seed, 77, 0.4))
    keep = torch.from_numpy(px.keep_mask(0x1234567890ABCDEF, 77, M, N, 0.4)).cuda()
    _close(D, (A.float() @ B.float().T) * keep, 3e-3, "dropout")
    frac = (keep == 0).float().mean().item()
    assert abs(frac - 102 / 256) < 0.01


@pytest.mark.parametrize("R,NO,KI", [(64, 128, 128), (3200, 256, 256), (1600, 1336, 668), (1000, 256, 512),
                                     (130, 200, 72), (51200, 256, 256)])
def test_gemm_wgrad(R, NO, KI):
    from multi_modal_foundation_model_b200 import ops
    dY, X = _mk(R, NO, seed=9, scale=0.1), _mk(R, KI, seed=10)
    dW = torch.ones(NO, KI, device="cuda")
    db = torch.full((NO,), 2.0, device="cuda")
    ops.gemm_wgrad(dY, X, dW, R=R, NO=NO, KI=KI, dbias=db)
    ref = dY.float().T @ X.float() + 1.0
    _close(dW, ref, 3e-3, f"wgrad {R}x{NO}x{KI}")
    _close(db, dY.float().sum(0) + 2.0, 3e-3, f"wgrad bias {R}x{NO}")
    dW2 = torch.zeros(NO, KI, device="cuda")
    ops.gemm_wgrad(dY, X, dW2, R=R, NO=NO, KI=KI)           # without the bias column
    _close(dW2, ref - 1.0, 3e-3, f"wgrad (no bias) {R}x{NO}x{KI}")


def test_colsum_and_cast():
    from multi_modal_foundation_model_b200 import ops
    dY = _mk(1000, 300, seed=11)
    out = torch.zeros(300, device="cuda")
    ops.colsum_bf16(dY, out, R=1000, NO=300)
    _close(out, dY.float().sum(0), 1e-4, "colsum")
    x = torch.randn(70, 50, device="cuda")
    y = torch.zeros(70, 56, device="cuda", dtype=torch.bfloat16)[:, :50]
    yt = torch.zeros(50, 72, device="cuda", dtype=torch.bfloat16)[:, :70]
    ops.cast_bf16(x, y, yt)
    assert torch.equal(y, x.to(torch.bfloat16)) and torch.equal(yt, x.to(torch.bfloat16).T)


@pytest.mark.parametrize("M,S,dropout", [(600, 200, True), (51200, 200, True), (1000, 100, False)])
def test_gemm_tn_rowdot_dropout_attention_prep(M, S, dropout):
    """Out-projection dgrad fused with the attention-backward preparation (MMFM_ACT_ROWDOT_DROP): per 32-column group
    (= head at d_head 32) rowdot[b, g, i] = <v, aux> and D = dropout(v) with the forward's output-dropout stream."""
    from _util import philox_keep_torch
    from multi_modal_foundation_model_b200 import ops
    from multi_modal_foundation_model_b200._lib import ACT_ROWDOT_DROP
    N, K = 256, 256
    A, B = _mk(M, K, seed=21, scale=0.3), _mk(N, K, seed=22, scale=0.1)
    O = _mk(M, N, seed=23)
    D = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16)
    Bt = M // S
    delta = torch.full((Bt, N // 32, S), 7.0, device="cuda")
    seed_val = 0x77AA55
    seed = torch.tensor([seed_val], dtype=torch.int64, device="cuda")
    ops.gemm_tn(A, B, D, act=ACT_ROWDOT_DROP, aux=O, rowdot=delta, rowdot_S=S,
                drop=ops.DropSpec(seed, 4098, 0.4) if dropout else ops.NO_DROP)
    v = A.float() @ B.float().T
    ref_delta = (v * O.float()).view(Bt, S, N // 32, 32).sum(-1).permute(0, 2, 1)
    _close(delta, ref_delta, 3e-3, "rowdot (delta)")
    if dropout:
        v = v * philox_keep_torch(seed_val, 4098, M, N, 0.4, "cuda")
    _close(D, v, 8e-3, "masked dO")
