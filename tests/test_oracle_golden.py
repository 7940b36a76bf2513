"""CPU: pins oracle/mm_oracle.py (the restatement every GPU parity test leans on) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py), and -- where /root/reference exists -- against the reference
run live."""
import numpy as np
import pytest
import torch

import _reference as ref
from _util import load_small, oracle_batch, oracle_params, rel_l2, small_config
from oracle import mm_oracle as orc


@pytest.mark.parametrize("mode", ["token_masking", "encoding", "decoding"])
def test_oracle_matches_golden(mode):
    z, W = load_small()
    spec = orc.OracleSpec.from_config(small_config(), ["ap", "behavior"])
    out, grads = orc.forward_backward(oracle_params(W), spec, oracle_batch(z, mode))
    assert abs(out.loss.item() - float(z[f"{mode}/loss"])) <= 1e-5 * abs(float(z[f"{mode}/loss"]))
    for m in ("ap", "behavior"):
        assert int(out.mod_n_examples[m]) == int(z[f"{mode}/n/{m}"])
        assert abs(out.mod_loss[m].item() - float(z[f"{mode}/mod_loss/{m}"])) <= 2e-5 * abs(float(z[f"{mode}/mod_loss/{m}"])) + 1e-4
        assert np.abs(out.mod_preds[m].detach().numpy() - z[f"{mode}/preds/{m}"]).max() < 2e-5
    for name, g in grads.items():
        if name.startswith("decoder_embeddings.") and name.endswith("mod_emb.weight"):
            continue            # alias of the encoder-side Parameter (mm.py:84-87); checked under that name
        if mode == "token_masking":
            gref = torch.from_numpy(z[f"{mode}/grad/{name}"])
            assert rel_l2(g, gref) < 2e-4 or gref.norm() < 1e-7, name
        else:
            gn = float(z[f"{mode}/gnorm/{name}"])
            assert abs(g.double().norm().item() - gn) <= 2e-4 * gn + 1e-8, name


def test_oracle_dropout_stream_statistics():
    from oracle import philox_ref as px
    m = px.keep_mask(1234, 5, 512, 256, 0.4)
    assert abs((m == 0).mean() - 102 / 256) < 5e-3
    assert abs(m.mean() - 1.0) < 1e-2                      # unbiased
    pm = px.prob_keep_mask(1234, 6, 512, 200, 0.4)
    assert pm.shape == (512, 200) and abs((pm == 0).mean() - 102 / 256) < 5e-3
    # different sites / seeds decorrelate
    m2 = px.keep_mask(1234, 6, 512, 256, 0.4)
    assert abs(((m == 0) & (m2 == 0)).mean() - (102 / 256) ** 2) < 5e-3
    # Philox4x32-10 known answer (Random123 kat: counter 0, key 0) pins the round function
    r = px.philox4x32(np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(1, np.uint32), np.zeros(1, np.uint32),
                      0, 0, rounds=10)
    assert [int(x[0]) for x in r] == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]


@pytest.mark.skipif(not ref.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_matches_reference_live():
    """Larger than the fixture: default 5+5-layer model, N=96, padded trials, all three modes."""
    from multi_modal_foundation_model_b200.synthetic import make_batch, make_mod_dict
    cfg = ref.load_config()
    torch.manual_seed(3)
    model = ref.build_reference_model(cfg, 96, 2).eval()
    W = {k: v.detach().clone() for k, v in model.state_dict().items()}
    spec = orc.OracleSpec.from_config(cfg["model"], ["ap", "behavior"])
    batch = make_batch(3, 96, 2, 100, step=1, pad_bins=10)
    _live_check(model, cfg, W, spec, batch)


def _live_check(model, cfg, W, spec, batch):
    from multi_modal_foundation_model_b200.synthetic import make_mod_dict
    for mode in ("token_masking", "decoding"):
        md = make_mod_dict(batch, ["ap", "behavior"], mode)
        model.zero_grad()
        o = model(md)
        o.loss.backward()
        ob = {m: dict(inputs=md[m]["inputs"], targets=md[m]["targets"], attn_mask=md[m]["inputs_attn_mask"],
                      timestamp=md[m]["inputs_timestamp"], mask=md[m]["inputs_mask"]) for m in ("ap", "behavior")}
        out, grads = orc.forward_backward(oracle_params(W), spec, ob)
        assert abs(out.loss.item() - o.loss.item()) < 1e-5 * abs(o.loss.item())
        for n, p in model.named_parameters():
            if p.grad is not None and p.grad.norm() > 1e-7:
                assert rel_l2(grads[n], p.grad) < 5e-4, n


@pytest.mark.skipif(not ref.available(), reason="reference tree not mounted")
@pytest.mark.parametrize("over", [{"decoder.decoder_causal_mask": True}, {"decoder.decoder_sep_mask": True},
                                  {"decoder.decoder_causal_mask": True, "decoder.decoder_sep_mask": True},
                                  {"encoder.transformer.use_scalenorm": True, "decoder.transformer.use_scalenorm": True}])
def test_oracle_causal_and_sep_branches_match_reference_live(over):
    """mm.py:178-194 (causal drops the key padding; sep ADDS cross-modality attention) and ScaleNorm (mm_utils.py:31-39):
    the oracle's restatement of these config branches against the reference run live (they are in no committed
    fixture)."""
    from multi_modal_foundation_model_b200.synthetic import make_batch
    cfg = ref.load_config(over)
    torch.manual_seed(4)
    model = ref.build_reference_model(cfg, 48, 2).eval()
    W = {k: v.detach().clone() for k, v in model.state_dict().items()}
    spec = orc.OracleSpec.from_config(cfg["model"], ["ap", "behavior"])
    _live_check(model, cfg, W, spec, make_batch(3, 48, 2, 100, step=5, pad_bins=15))


@pytest.mark.skipif(not ref.available(), reason="reference tree not mounted")
def test_oracle_cross_entropy_extension_matches_reparameterised_reference():
    """Choice / block streams + CE (BASELINE.json north_star; the reference has neither): the
    oracle's 'ce' branch == the UNMODIFIED reference classes re-parameterised -- two extra one-hot modalities through the
    ordinary embedders, ``loss_mod[mod] = TokenCrossEntropy()`` (dict assignment, class code unchanged)."""
    from multi_modal_foundation_model_b200.losses import TokenCrossEntropy, one_hot_stream
    from multi_modal_foundation_model_b200.synthetic import make_batch
    mods = ["ap", "behavior", "choice", "block"]
    K = {"choice": 2, "block": 3}
    cfg = ref.load_config({"encoder.embedder.n_modality": 4, "decoder.embedder.n_modality": 4,
                           "encoder.transformer.n_layers": 2, "decoder.transformer.n_layers": 2})
    ref.activate()
    from multi_modal.mm import MultiModal
    from multi_modal.encoder_embeddings import EncoderEmbedding
    from multi_modal.decoder_embeddings import DecoderEmbedding
    torch.manual_seed(8)
    B, T, N = 4, 100, 40
    chan = {"ap": N, "behavior": 2, **K}
    enc = {m: EncoderEmbedding(hidden_size=256, n_channel=chan[m], config=cfg.model.encoder) for m in mods}
    dec = {m: DecoderEmbedding(hidden_size=256, n_channel=chan[m], output_channel=chan[m], config=cfg.model.decoder) for m in mods}
    model = MultiModal(enc, dec, avail_mod=mods, config=cfg.model, share_modality_embeddings=True).eval()
    for m in K:
        model.loss_mod[m] = TokenCrossEntropy()
    W = {k: v.detach().clone() for k, v in model.state_dict().items()}
    batch = make_batch(B, N, 2, T, step=6, pad_bins=10)
    g = torch.Generator().manual_seed(2)
    xs = {"ap": batch["spikes_data"], "behavior": batch["target"]}
    for m, k in K.items():
        xs[m] = one_hot_stream(torch.randint(0, k, (B, 1), generator=g).expand(B, T), k)
    masks = {m: (torch.rand(B, T, generator=g) < 0.4).long() for m in mods}
    md = {}
    for i, m in enumerate(mods):
        md[m] = dict(inputs=xs[m].clone(), targets=xs[m].clone(), inputs_attn_mask=batch["time_attn_mask"],
                     inputs_timestamp=batch["spikes_timestamps"], inputs_modality=torch.tensor(i), masking_mode=None,
                     eval_mask=masks[m][:, :, None].expand(B, T, chan[m]).contiguous(),
                     inputs_regions=np.array([["CA1"] * N] * B))
    o = model(md)
    o.loss.backward()
    spec = orc.OracleSpec.from_config(cfg["model"], mods)
    spec.loss_kind.update({m: "ce" for m in K})
    ob = {m: dict(inputs=xs[m], targets=xs[m], attn_mask=batch["time_attn_mask"], timestamp=batch["spikes_timestamps"],
                  mask=masks[m] & batch["time_attn_mask"]) for m in mods}
    out, grads = orc.forward_backward(oracle_params(W), spec, ob)
    assert abs(out.loss.item() - o.loss.item()) < 1e-5 * abs(o.loss.item())
    for m in mods:
        assert int(out.mod_n_examples[m]) == int(o.mod_n_examples[m])
        assert abs(out.mod_loss[m].item() - o.mod_loss[m].item()) < 2e-5 * abs(o.mod_loss[m].item()) + 1e-5
    for n, p in model.named_parameters():
        if p.grad is not None and p.grad.norm() > 1e-7:
            assert rel_l2(grads[n], p.grad) < 5e-4, n
