"""GPU parity of the train_baseline.py linear models against the oracle restatement
(oracle/mm_oracle.py: baseline_decoder / baseline_encoder) -- config 1 of BASELINE.json: B=16, T=100, N=512."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _grads(loss, params):
    return torch.autograd.grad(loss, params)


def test_baseline_decoder_matches_oracle():
    from multi_modal_foundation_model_b200.baselines import BaselineDecoder
    from multi_modal_foundation_model_b200.synthetic import make_batch
    from oracle import mm_oracle as orc
    torch.manual_seed(0)
    m = BaselineDecoder(512, 2).cuda()
    b = make_batch(16, 512, 2, 100)
    x, y = b["spikes_data"].cuda(), b["target"].cuda()
    out = m({"inputs": x, "targets": y})
    out.loss.backward()
    P = {"layer.weight": m.layer.weight.detach().cpu().clone().requires_grad_(True),
         "layer.bias": m.layer.bias.detach().cpu().clone().requires_grad_(True)}
    ref_loss, ref_preds = orc.baseline_decoder(P, x.cpu(), y.cpu())
    gw, gb = _grads(ref_loss, [P["layer.weight"], P["layer.bias"]])
    assert out.n_examples == 16 and out.preds.shape == (16, 100, 2)
    assert abs(out.loss.item() - ref_loss.item()) <= 2e-3 * abs(ref_loss.item())
    assert (out.preds.cpu() - ref_preds).abs().max().item() < 3e-2
    for got, ref in ((m.layer.weight.grad.cpu(), gw), (m.layer.bias.grad.cpu(), gb)):
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 2e-2, rel


def test_baseline_encoder_matches_oracle():
    from multi_modal_foundation_model_b200.baselines import BaselineEncoder
    from multi_modal_foundation_model_b200.synthetic import make_batch
    from oracle import mm_oracle as orc
    torch.manual_seed(0)
    N = 128
    m = BaselineEncoder(2, N, seq_len=100).cuda()
    with torch.no_grad():
        m.layer.weight.mul_(0.1)
    b = make_batch(16, N, 2, 100)
    x, y = b["target"].cuda(), b["spikes_data"].cuda()
    out = m({"inputs": x, "targets": y})
    out.loss.backward()
    P = {"layer.weight": m.layer.weight.detach().cpu().clone().requires_grad_(True),
         "layer.bias": m.layer.bias.detach().cpu().clone().requires_grad_(True)}
    ref_loss, ref_preds = orc.baseline_encoder(P, x.cpu(), y.cpu())
    gw, gb = _grads(ref_loss, [P["layer.weight"], P["layer.bias"]])
    assert abs(out.loss.item() - ref_loss.item()) <= 2e-3 * abs(ref_loss.item())
    assert (out.preds.cpu() - ref_preds).abs().max().item() < 3e-2
    for got, ref in ((m.layer.weight.grad.cpu(), gw), (m.layer.bias.grad.cpu(), gb)):
        rel = ((got - ref).norm() / ref.norm()).item()
        assert rel < 2e-2, rel
