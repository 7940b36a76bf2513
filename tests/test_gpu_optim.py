"""Fused AdamW (optim.AdamW -> mmfm_adamw_step) against torch.optim.AdamW on the same parameters and gradients
(reference train_multi_modal.py:197-210: AdamW lr 1e-4 wd 0.01 eps 1e-8 driven by OneCycleLR)."""
import pytest
import torch

from _util import small_config

pytestmark = pytest.mark.gpu


def test_adamw_matches_torch_over_steps():
    from multi_modal_foundation_model_b200.model import build_model
    from multi_modal_foundation_model_b200.optim import AdamW
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    torch.manual_seed(1)
    model = build_model(40, 2, small_config()).cuda().eval()
    opt = AdamW(model.parameters(), lr=1e-3, weight_decay=0.01, eps=1e-8)
    sched = torch.optim.lr_scheduler.OneCycleLR(opt, total_steps=10, max_lr=1e-3, pct_start=0.3, div_factor=10)
    # shadow copy driven by torch's own optimizer, fed with the gradients the engine produced
    ref = {n: p.detach().clone() for n, p in model.named_parameters()}
    ref_params = [torch.nn.Parameter(v) for v in ref.values()]
    ropt = torch.optim.AdamW(ref_params, lr=1e-3, weight_decay=0.01, eps=1e-8)
    rsched = torch.optim.lr_scheduler.OneCycleLR(ropt, total_steps=10, max_lr=1e-3, pct_start=0.3, div_factor=10)
    losses = []
    for step in range(4):
        batch = make_batch(3, 40, 2, 100, step=step)
        md = make_mod_dict(batch, ["ap", "behavior"], ("encoding", "decoding", "token_masking")[step % 3], device="cuda")
        torch.manual_seed(100 + step)
        out = model(md)
        out.loss.backward()
        losses.append(out.loss.item())
        for rp, (n, p) in zip(ref_params, model.named_parameters()):
            rp.grad = p.grad.detach().clone()
        opt.step()
        sched.step()
        ropt.step()
        rsched.step()
        opt.zero_grad(set_to_none=True)
        for rp, (n, p) in zip(ref_params, model.named_parameters()):
            assert torch.allclose(p.detach(), rp.detach(), rtol=2e-6, atol=2e-8), (step, n, (p - rp).abs().max().item())
    # parameters really moved and the next forward pass reads the updated weights (bf16 shadows refreshed)
    assert losses[3] != losses[0]
    sd = opt.state_dict()
    assert len(sd["state"]) == len(list(model.parameters()))
    name0, p0 = next(iter(model.named_parameters()))
    assert torch.equal(opt.state[p0]["exp_avg"], ropt.state[ref_params[0]]["exp_avg"]) or torch.allclose(
        opt.state[p0]["exp_avg"], ropt.state[ref_params[0]]["exp_avg"], rtol=1e-5, atol=1e-9)


def test_adamw_requires_engine_parameters():
    from multi_modal_foundation_model_b200._lib import MmfmError
    from multi_modal_foundation_model_b200.optim import AdamW
    w = torch.nn.Parameter(torch.zeros(8, device="cuda"))
    w.grad = torch.ones_like(w)
    with pytest.raises(MmfmError):
        AdamW([w]).step()


def test_adamw_leaves_parameters_without_gradient_untouched():
    """torch.optim.AdamW skips a parameter whose .grad is None (no update, no weight decay, no moment decay); the fused
    step covers the whole flat buffer in one launch and must put such ranges back."""
    from multi_modal_foundation_model_b200.model import build_model
    from multi_modal_foundation_model_b200.optim import AdamW
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    torch.manual_seed(2)
    model = build_model(40, 2, small_config()).cuda().eval()
    opt = AdamW(model.parameters(), lr=1e-2, weight_decay=0.1, eps=1e-8)
    md = lambda: make_mod_dict(make_batch(3, 40, 2, 100, step=0), ["ap", "behavior"], "encoding", device="cuda")
    model(md()).loss.backward()
    opt.step()                                        # builds the moments
    model(md()).loss.backward()
    name, frozen = "encoder.0.mlp.up_proj.weight", dict(model.named_parameters())["encoder.0.mlp.up_proj.weight"]
    w0 = frozen.detach().clone()
    m0 = opt.state[frozen]["exp_avg"].clone()
    other = dict(model.named_parameters())["encoder.0.mlp.down_proj.weight"]
    o0 = other.detach().clone()
    frozen.grad = None
    opt.step()
    assert torch.equal(frozen.detach(), w0) and torch.equal(opt.state[frozen]["exp_avg"], m0), name
    assert not torch.equal(other.detach(), o0)
