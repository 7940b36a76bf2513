"""GPU parity of the bandwidth kernels (csrc/norm.cu, csrc/glue.cu) against plain fp32 torch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("R,H", [(3200, 256), (1000, 1024), (77, 128), (300, 512), (50, 200)])
def test_layernorm_fwd(R, H):
    from multi_modal_foundation_model_b200 import ops
    torch.manual_seed(R + H)
    x = torch.randn(R, H, device="cuda") * 2 + 0.5
    g, b = torch.randn(H, device="cuda"), torch.randn(H, device="cuda")
    y = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(R, device="cuda"), torch.empty(R, device="cuda")
    ops.layernorm_fwd(x, g, b, y, mean, rstd, R=R, H=H)
    ref = F.layer_norm(x, (H,), g, b, 1e-5)
    # bf16 output: one rounding of the fp32 result (2^-8 relative) plus the fp32 evaluation-order difference
    assert ((y.float() - ref).abs() <= ref.abs() * 2.0 ** -7 + 2e-3).all()
    assert (mean - x.mean(1)).abs().max().item() < 1e-5
    assert ((rstd - (x.var(1, unbiased=False) + 1e-5).rsqrt()).abs() / rstd).max().item() < 1e-4


def test_layernorm_fwd_modmajor():
    from multi_modal_foundation_model_b200 import ops
    B, T, M, H = 3, 100, 2, 256
    S = T * M
    x = torch.randn(B * S, H, device="cuda")
    g, b = torch.ones(H, device="cuda"), torch.zeros(H, device="cuda")
    y = torch.empty(B * S, H, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(B * S, device="cuda"), torch.empty(B * S, device="cuda")
    ops.layernorm_fwd(x, g, b, y, mean, rstd, R=B * S, H=H, modmajor_T=T, S=S)
    ref = F.layer_norm(x, (H,)).view(B, M, T, H).transpose(0, 1).reshape(B * S, H)
    assert (y.float() - ref).abs().max().item() < 4e-2


@pytest.mark.parametrize("R,H,mm", [(3200, 256, False), (500, 1024, False), (600, 256, True), (123, 128, False)])
def test_layernorm_bwd(R, H, mm):
    from multi_modal_foundation_model_b200 import ops
    T, S = 100, 200
    x = (torch.randn(R, H, device="cuda") * 1.5).requires_grad_(True)
    g = torch.randn(H, device="cuda", requires_grad=True)
    b = torch.randn(H, device="cuda", requires_grad=True)
    dy = (torch.randn(R, H, device="cuda") * 0.1).to(torch.bfloat16)
    dres = torch.randn(R, H, device="cuda") * 0.1
    y = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(R, device="cuda"), torch.empty(R, device="cuda")
    ops.layernorm_fwd(x.detach(), g.detach(), b.detach(), y, mean, rstd, R=R, H=H)
    ref = F.layer_norm(x, (H,), g, b, 1e-5)
    dy_nat = dy.float()
    if mm:
        B = R // S
        dy_nat = dy.float().view(S // T, B, T, H).transpose(0, 1).reshape(R, H)
    ref.backward(dy_nat)
    dx = torch.empty(R, H, device="cuda")
    dxb = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    dg, db = torch.zeros(H, device="cuda"), torch.zeros(H, device="cuda")
    ops.layernorm_bwd(dy, x.detach(), mean, rstd, g.detach(), dres, dx, dxb, ops.NO_DROP, dg, db, R=R, H=H,
                      modmajor_T=T if mm else 0, S=S if mm else 0)
    assert (dx - (x.grad + dres)).abs().max().item() < 2e-4
    assert (dxb.float() - dx).abs().max().item() < 1e-2
    assert (dg - g.grad).abs().max().item() < 2e-3 * max(1.0, g.grad.abs().max().item())
    assert (db - b.grad).abs().max().item() < 2e-3 * max(1.0, b.grad.abs().max().item())


def test_layernorm_bwd_dropout_copy():
    from multi_modal_foundation_model_b200 import ops
    from oracle import philox_ref as px
    R, H = 400, 256
    x = torch.randn(R, H, device="cuda")
    g = torch.ones(H, device="cuda")
    dy = torch.randn(R, H, device="cuda").to(torch.bfloat16)
    y = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    mean, rstd = torch.empty(R, device="cuda"), torch.empty(R, device="cuda")
    ops.layernorm_fwd(x, g, torch.zeros(H, device="cuda"), y, mean, rstd, R=R, H=H)
    dx = torch.empty(R, H, device="cuda")
    dxb = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    seed = torch.tensor([99], dtype=torch.int64, device="cuda")
    ops.layernorm_bwd(dy, x, mean, rstd, g, None, dx, dxb, ops.DropSpec(seed, 5, 0.4), None, None, R=R, H=H)
    keep = torch.from_numpy(px.keep_mask(99, 5, R, H, 0.4)).cuda()
    ref = dx * keep
    assert ((dxb.float() - ref).abs() <= 2.0 ** -7 * ref.abs() + 1e-6).all()   # one bf16 rounding


def test_mask_prep_and_embed():
    from multi_modal_foundation_model_b200 import ops
    B, T, H = 5, 100, 256
    S = 2 * T
    g = torch.Generator(device="cuda").manual_seed(3)
    big = (torch.rand(B, T, 7, generator=g, device="cuda") < 0.3).long()     # eval_mask-like (B,T,C)
    m1 = (torch.rand(B, T, generator=g, device="cuda") < 0.3).long()
    attn = torch.ones(B, T, dtype=torch.int64, device="cuda")
    attn[:, 80:] = 0
    zero = torch.zeros(S, dtype=torch.uint8, device="cuda")
    kval = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
    tmask = torch.zeros(B, S, dtype=torch.uint8, device="cuda")
    nex = torch.zeros(2, dtype=torch.int64, device="cuda")
    inv = torch.zeros(1, device="cuda")
    ops.mask_prep([big[:, :, 0], m1], [attn, attn], [512, 2], zero, kval, tmask, nex, inv)
    mk = torch.cat([big[:, :, 0] & attn, m1 & attn], 1)
    assert torch.equal(tmask.long(), mk)
    assert torch.equal(zero.long(), (mk[0] == 1).long())
    assert torch.equal(kval.long(), torch.cat([attn, attn], 1))
    n_ref = torch.stack([mk[:, :T].sum() * 512, mk[:, T:].sum() * 2])
    assert torch.equal(nex, n_ref)
    assert abs(inv.item() * n_ref.sum().item() - 1.0) < 1e-6
    # embedding assemble + backward
    mod_emb = torch.randn(2, H, device="cuda")
    pos = torch.randn(T, H, device="cuda")
    ts = torch.arange(T, device="cuda")[None].expand(B, T).contiguous()
    ts[1] = torch.randint(0, T, (T,), generator=g, device="cuda")
    emb = torch.zeros(B * S, H, device="cuda")
    ops.embed_assemble(mod_emb[1], pos, ts, emb, B=B, T=T, S=S, off=T, H=H)
    ref = mod_emb[1][None, None] + pos[ts]
    assert torch.equal(emb.view(B, S, H)[:, T:], ref)
    gr, g2 = torch.randn(B * S, H, device="cuda"), torch.randn(B * S, H, device="cuda")
    dpos, dmod = torch.zeros(T, H, device="cuda"), torch.zeros(H, device="cuda")
    ops.embed_assemble_bwd(gr, g2, ts, dpos, dmod, B=B, T=T, S=S, off=T, H=H)
    gsum = (gr + g2).view(B, S, H)[:, T:]
    dpos_ref = torch.zeros(T, H, device="cuda").index_add_(0, ts.reshape(-1), gsum.reshape(-1, H))
    assert (dpos - dpos_ref).abs().max().item() < 1e-4
    assert (dmod - gsum.sum((0, 1))).abs().max().item() < 1e-3


@pytest.mark.parametrize("C", [1, 2, 5])
def test_smallc_embed_and_head(C):
    from multi_modal_foundation_model_b200 import ops
    from multi_modal_foundation_model_b200._lib import ACT_SOFTSIGN
    B, T, H = 6, 100, 256
    S, off = 2 * T, T
    dev = "cuda"
    inp = torch.randn(B, T, C, device=dev)
    W1 = torch.randn(2 * C, C, device=dev, requires_grad=True)
    b1 = torch.randn(2 * C, device=dev, requires_grad=True)
    W2 = (torch.randn(H, 2 * C, device=dev) * 0.3).requires_grad_(True)
    b2 = torch.randn(H, device=dev, requires_grad=True)
    emb = torch.randn(B * S, H, device=dev)
    zero = torch.zeros(S, dtype=torch.uint8, device=dev)
    zero[off + 7] = 1
    x = torch.zeros(B * S, H, device=dev)
    hid = torch.zeros(B * T, 2 * C, device=dev)
    ops.smallc_embed_fwd(inp, W1, b1, W2, b2, emb, x, hid, zero, ops.NO_DROP, 1.0, ACT_SOFTSIGN, B=B, T=T, S=S, off=off,
                         Cc=C, H=H)
    h_ref = F.softsign(F.linear(inp, W1, b1))
    tok = F.linear(h_ref, W2, b2)
    keep = torch.ones(T, device=dev)
    keep[7] = 0
    tok = tok * keep[None, :, None]
    ref = tok + emb.view(B, S, H)[:, off:]
    assert (x.view(B, S, H)[:, off:] - ref).abs().max().item() < 1e-4
    dx = torch.randn(B * S, H, device=dev)
    ref.backward(dx.view(B, S, H)[:, off:])
    dW1, db1 = torch.zeros_like(W1), torch.zeros_like(b1)
    dW2, db2 = torch.zeros_like(W2), torch.zeros_like(b2)
    ops.smallc_embed_bwd(inp, hid, W2.detach(), dx, zero, ops.NO_DROP, 1.0, ACT_SOFTSIGN, dW1, db1, dW2, db2, B=B, T=T,
                         S=S, off=off, Cc=C, H=H)
    for nm, got, rf in (("dW1", dW1, W1.grad), ("db1", db1, b1.grad), ("dW2", dW2, W2.grad), ("db2", db2, b2.grad)):
        sc = rf.abs().max().item() + 1e-6
        assert (got - rf).abs().max().item() < 2e-3 * sc, nm
    # head
    R = B * T
    y = torch.randn(R, H, device=dev).to(torch.bfloat16)
    Wo = torch.randn(C, H, device=dev, requires_grad=True)
    bo = torch.randn(C, device=dev, requires_grad=True)
    preds = torch.zeros(R, C, device=dev)
    ops.smallc_head_fwd(y, Wo, bo, preds, R=R, H=H, Cc=C)
    yf = y.float().requires_grad_(True)
    pref = F.linear(yf, Wo, bo)
    assert (preds - pref).abs().max().item() < 1e-3
    dp = torch.zeros(R, 8, device=dev, dtype=torch.bfloat16)
    dp[:, :C] = torch.randn(R, C, device=dev).to(torch.bfloat16)
    pref.backward(dp[:, :C].float())
    dy = torch.zeros(R, H, device=dev, dtype=torch.bfloat16)
    dWo, dbo = torch.zeros_like(Wo), torch.zeros_like(bo)
    ops.smallc_head_bwd(y, Wo.detach(), dp, dy, dWo, dbo, R=R, H=H, Cc=C)
    assert (dy.float() - yf.grad).abs().max().item() < 3e-2 * (yf.grad.abs().max().item() + 1e-6)
    assert (dWo - Wo.grad).abs().max().item() < 2e-3 * Wo.grad.abs().max().item()
    assert (dbo - bo.grad).abs().max().item() < 2e-3 * bo.grad.abs().max().item()


@pytest.mark.parametrize("kind,C", [(0, 668), (0, 512), (1, 2), (1, 1), (0, 333)])
def test_loss_fwd_bwd(kind, C):
    from multi_modal_foundation_model_b200 import ops
    B, T = 7, 100
    S, off = 2 * T, T if kind == 1 else 0
    dev = "cuda"
    preds = (torch.randn(B * T, C, device=dev) * 0.5).requires_grad_(True)
    tg = torch.poisson(torch.full((B * T, C), 0.3, device=dev)) if kind == 0 else torch.randn(B * T, C, device=dev)
    tmask = (torch.rand(B, S, device=dev) < 0.3).to(torch.uint8)
    w = tmask[:, off:off + T].reshape(B * T, 1).float()
    ell = (torch.exp(preds) - tg * preds) if kind == 0 else (preds - tg) ** 2
    n = (w.sum() * C)
    inv_n = (1.0 / n).reshape(1).float()
    loss_ref = (ell * w).sum()
    (loss_ref / n).backward()
    npart = 64
    partials = torch.zeros(npart, device=dev)
    Cp = (C + 7) // 8 * 8
    dp = torch.zeros(B * T, Cp, device=dev, dtype=torch.bfloat16)
    ops.loss_fwd_bwd(preds.detach(), tg, tmask, inv_n, kind, partials, dp, B=B, T=T, Cc=C, S=S, off=off)
    mod_loss, loss = torch.zeros(1, device=dev), torch.zeros(1, device=dev)
    ops.loss_finalize(partials, npart, 1, inv_n, mod_loss, loss)
    assert abs(mod_loss.item() - loss_ref.item()) < 1e-4 * abs(loss_ref.item())
    assert abs(loss.item() - (loss_ref / n).item()) < 1e-4 * abs((loss_ref / n).item())
    gmax = preds.grad.abs().max().item()
    assert (dp[:, :C].float() - preds.grad).abs().max().item() < 1e-2 * gmax


def test_cast_multi_and_scale():
    import ctypes as Cc
    from multi_modal_foundation_model_b200 import ops
    from multi_modal_foundation_model_b200._lib import CastItem
    a = torch.randn(100, 70, device="cuda")
    b = torch.randn(33, 256, device="cuda")
    ya = torch.zeros(100, 72, device="cuda", dtype=torch.bfloat16)
    yat = torch.zeros(70, 104, device="cuda", dtype=torch.bfloat16)
    ybt = torch.zeros(256, 40, device="cuda", dtype=torch.bfloat16)
    items = (CastItem * 2)()
    items[0] = CastItem(a.data_ptr(), 70, ya.data_ptr(), 72, yat.data_ptr(), 104, 100, 70, 0, 0)
    t0 = ((100 + 31) // 32) * ((70 + 31) // 32)
    items[1] = CastItem(b.data_ptr(), 256, None, 0, ybt.data_ptr(), 40, 33, 256, t0, 0)
    total = t0 + ((33 + 31) // 32) * (256 // 32)
    raw = torch.frombuffer(bytearray(bytes(items)), dtype=torch.uint8).cuda()
    ops.cast_bf16_multi(raw, 2, total)
    assert torch.equal(ya[:, :70], a.to(torch.bfloat16))
    assert torch.equal(yat[:, :100], a.to(torch.bfloat16).T)
    assert torch.equal(ybt[:, :33], b.to(torch.bfloat16).T)
    x = torch.randn(1003, device="cuda")
    x0 = x.clone()
    ops.scale_inplace(x, torch.tensor([1.0], device="cuda"))
    assert torch.equal(x, x0)
    ops.scale_inplace(x, torch.tensor([0.5], device="cuda"))
    assert torch.equal(x, x0 * 0.5)


@pytest.mark.parametrize("R,H,n", [(1000, 256, 5), (77, 128, 8), (300, 512, 3), (64, 256, 1)])
def test_layernorm_multi_matches_n_separate_layernorms(R, H, n):
    """The decoder layers' context_norm share their input (decoder_embeddings.py:141-145): one launch per direction
    against n torch LayerNorms over the same tensor and the sum of their input gradients."""
    from multi_modal_foundation_model_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    x = (torch.randn(R, H, generator=g, device="cuda") * 1.7 + 0.3).requires_grad_(True)
    gam = [(1.0 + 0.2 * torch.randn(H, generator=g, device="cuda")).requires_grad_(True) for _ in range(n)]
    bet = [(0.1 * torch.randn(H, generator=g, device="cuda")).requires_grad_(True) for _ in range(n)]
    ys = [torch.empty(R, H, device="cuda", dtype=torch.bfloat16) for _ in range(n)]
    mean, rstd = torch.empty(R, device="cuda"), torch.empty(R, device="cuda")
    ops.layernorm_fwd_multi(x.detach(), [t.detach() for t in gam], [t.detach() for t in bet], ys, mean, rstd, R=R, H=H)
    refs = [torch.nn.functional.layer_norm(x, (H,), gam[l], bet[l], 1e-5) for l in range(n)]
    for l in range(n):
        assert (ys[l].float() - refs[l]).abs().max().item() < 3e-2
    assert torch.allclose(mean, x.detach().mean(1), atol=1e-5)
    dys = [(torch.randn(R, H, generator=g, device="cuda") * 0.1).to(torch.bfloat16) for _ in range(n)]
    sum(( refs[l] * dys[l].float()).sum() for l in range(n)).backward()
    dx = torch.empty(R, H, device="cuda")
    dxb = torch.empty(R, H, device="cuda", dtype=torch.bfloat16)
    dgs = [torch.zeros(H, device="cuda") for _ in range(n)]
    dbs = [torch.zeros(H, device="cuda") for _ in range(n)]
    ops.layernorm_bwd_multi(dys, x.detach(), mean, rstd, [t.detach() for t in gam], dx, dxb, dgs, dbs, R=R, H=H)
    torch.cuda.synchronize()
    scale = x.grad.abs().max().item()
    assert (dx - x.grad).abs().max().item() < 2e-3 * scale + 1e-5
    assert (dxb.float() - x.grad).abs().max().item() < 1e-2 * scale + 1e-4
    for l in range(n):
        assert (dgs[l] - gam[l].grad).abs().max().item() < 2e-3 * gam[l].grad.abs().max().item() + 1e-4
        assert (dbs[l] - bet[l].grad).abs().max().item() < 2e-3 * bet[l].grad.abs().max().item() + 1e-4
