"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32/fp64, autograd for gradients) of the
reference's masked multi-modal encoder/decoder forward pass and loss.

Nothing in the product package imports this file.  It is the checker used by ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the UNMODIFIED reference classes
(imported read-only from ``/root/reference/src`` in the build container) and stores inputs, weights,
loss, predictions and gradients under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this
restatement against those fixtures (and live against the reference when ``/root/reference`` exists).
The reference itself ships no tests or golden vectors.

Every function cites the reference lines it restates (paths relative to ``/root/reference``).
The restatement is functional: weights come from a ``state_dict``-style mapping with the reference's
key names, tokens of all modalities live in one packed (B, S=M*T, H) layout and the (B,S,S) int64
attention masks of the reference are evaluated as predicates instead of being materialised.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Mapping, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import philox_ref as px


@dataclass
class OracleSpec:
    """The handful of hyper-parameters the forward pass reads (``mm.yaml``)."""
    avail_mod: List[str]
    n_enc_layers: int
    n_dec_layers: int
    hidden: int
    n_heads: int
    embed_scale: float = 1.0          # embedder.scale (sqrt(H) when null; encoder_embeddings.py:34)
    embed_act: str = "softsign"
    embed_dropout: float = 0.2
    dropout: float = 0.4
    decoder_sep_mask: bool = False
    decoder_causal_mask: bool = False
    loss_kind: Dict[str, str] = field(default_factory=lambda: {"ap": "poisson", "behavior": "mse"})

    @staticmethod
    def from_config(cfg, avail_mod: Sequence[str]) -> "OracleSpec":
        et, emb = cfg["encoder"]["transformer"], cfg["encoder"]["embedder"]
        H = et["hidden_size"]
        scale = emb["scale"]
        return OracleSpec(
            avail_mod=list(avail_mod),
            n_enc_layers=et["n_layers"], n_dec_layers=cfg["decoder"]["transformer"]["n_layers"],
            hidden=H, n_heads=et["n_heads"],
            embed_scale=float(H ** 0.5 if scale is None else scale),
            embed_act=emb["act"], embed_dropout=float(emb["dropout"]), dropout=float(et["dropout"]),
            decoder_sep_mask=bool(cfg["decoder"]["decoder_sep_mask"]),
            decoder_causal_mask=bool(cfg["decoder"]["decoder_causal_mask"]),
            loss_kind={m: ("poisson" if m == "ap" else "mse") for m in avail_mod},
        )


@dataclass
class OracleOut:
    loss: torch.Tensor
    mod_loss: Dict[str, torch.Tensor]
    mod_n_examples: Dict[str, torch.Tensor]
    mod_preds: Dict[str, torch.Tensor]
    mod_targets: Dict[str, torch.Tensor]
    taps: Dict[str, torch.Tensor]


class _Drop:
    """Dropout with the CUDA path's own Philox stream (see ``philox_ref``); identity when off."""

    def __init__(self, seed: Optional[int]):
        self.seed = seed

    def __call__(self, x: torch.Tensor, p: float, site: int, prob_layout: bool = False) -> torch.Tensor:
        if self.seed is None or p <= 0.0:
            return x
        cols = x.shape[-1]
        rows = x.numel() // cols
        fn = px.prob_keep_mask if prob_layout else px.keep_mask
        m = torch.from_numpy(fn(self.seed, site, rows, cols, p)).to(x.dtype)
        return x * m.reshape(x.shape)


def _linear(P: Mapping[str, torch.Tensor], name: str, x: torch.Tensor) -> torch.Tensor:
    b = P.get(name + ".bias")
    return F.linear(x, P[name + ".weight"], b)


def _layer_norm(P, name: str, x: torch.Tensor) -> torch.Tensor:
    if name + ".scale" in P:
        # ScaleNorm (mm_utils.py:31-39, use_scalenorm): x * scale / max(||x||, eps), eps 1e-5
        return x * (P[name + ".scale"] / torch.norm(x, dim=-1, keepdim=True).clamp(min=1e-5))
    # nn.LayerNorm(H), eps 1e-5, affine (encoder_embeddings.py:98,100; mm.py:72,77)
    return F.layer_norm(x, (x.shape[-1],), P[name + ".weight"], P[name + ".bias"], 1e-5)


def embed_tokens(P, prefix: str, spec: OracleSpec, mod_index: int, inputs, timestamps, drop: _Drop,
                 site: int):
    """encoder_embeddings.py:44-61 == decoder_embeddings.py:43-61 (own weights per side).

    returns tokens (B,T,H) [after embedding dropout] and emb (B,T,H) = mod_emb[m] + pos_embed[ts]."""
    h = _linear(P, prefix + ".token_embed", inputs)
    if spec.embed_act == "softsign":
        h = h / (1.0 + h.abs())
    elif spec.embed_act != "identity":
        raise NotImplementedError(spec.embed_act)
    h = h * spec.embed_scale
    tok = _linear(P, prefix + ".projection", h)
    emb = P[prefix + ".mod_emb.weight"][mod_index][None, None, :].expand(tok.shape).clone()
    pos_key = prefix + ".pos_embed.weight"
    if pos_key in P:
        emb = emb + P[pos_key][timestamps]
    return drop(tok, spec.embed_dropout, site), emb


def attention(P, prefix: str, spec: OracleSpec, x_q, x_kv, allowed, drop: _Drop, site_prob: int,
              site_out: int):
    """mm_utils.py:97-114 (self) and :139-152 (cross): q/k/v projections, softmax(QK^T/sqrt(d) with
    -inf where not allowed) [prob dropout] V, [output dropout], out_proj.  ``allowed`` is the bool
    (B,Sq,Sk) predicate with SDPA's True = attend convention."""
    B, Sq, H = x_q.shape
    Sk = x_kv.shape[1]
    nh = spec.n_heads
    d = H // nh
    q = _linear(P, prefix + ".query", x_q).view(B, Sq, nh, d).transpose(1, 2)
    k = _linear(P, prefix + ".key", x_kv).view(B, Sk, nh, d).transpose(1, 2)
    v = _linear(P, prefix + ".value", x_kv).view(B, Sk, nh, d).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(d))
    s = s.masked_fill(~allowed[:, None, :, :], float("-inf"))
    p = torch.softmax(s, dim=-1)
    p = drop(p, spec.dropout, site_prob, prob_layout=True)   # rows = ((b*nh+h)*Sq+i), cols = Sk
    o = (p @ v).transpose(1, 2).reshape(B, Sq, H)
    o = drop(o, spec.dropout, site_out)
    return _linear(P, prefix + ".out_proj", o)


def mlp(P, prefix: str, spec: OracleSpec, x, drop: _Drop, site: int):
    """mm_utils.py:50-52: down(gelu_erf(up(x))) then dropout."""
    u = _linear(P, prefix + ".up_proj", x)
    g = F.gelu(u)  # ACT2FN['gelu'] == exact erf GELU
    return drop(_linear(P, prefix + ".down_proj", g), spec.dropout, site)


def encoder_allowed(attn_cat: torch.Tensor) -> torch.Tensor:
    """mm.py:152-158: eye(S) | (ones & key_valid[b, j])."""
    B, S = attn_cat.shape
    eye = torch.eye(S, dtype=torch.bool)
    return eye[None] | attn_cat.bool()[:, None, :].expand(B, S, S)


def decoder_allowed(attn_cat: torch.Tensor, mod_ids: torch.Tensor, spec: OracleSpec) -> torch.Tensor:
    """mm.py:178-194: key padding, or lower-triangular (padding dropped) when causal; OR'ed with
    'different modality' when ``decoder_sep_mask`` (which, under SDPA's convention, ADDS attention)."""
    B, S = attn_cat.shape
    if spec.decoder_causal_mask:
        a = torch.ones(S, S).tril().bool()[None].expand(B, S, S)  # create_context_mask(0,-1,S)
    else:
        a = attn_cat.bool()[:, None, :].expand(B, S, S)
    if spec.decoder_sep_mask:
        a = a | (mod_ids[None, None, :] != mod_ids[None, :, None])
    return a


def forward(P: Mapping[str, torch.Tensor], spec: OracleSpec, batch: Mapping[str, Mapping[str, torch.Tensor]],
            dropout_seed: Optional[int] = None, keep_taps: bool = False) -> OracleOut:
    """mm.py:242-308 from the point where the per-modality (B,T) masks are known.

    ``batch[mod]`` holds ``inputs`` (B,T,C), ``targets`` (B,T,C), ``attn_mask`` (B,T) int64,
    ``timestamp`` (B,T) int64 and ``mask`` (B,T) int64 (= ``mask[:,:,0] & inputs_attn_mask``,
    mm.py:270).  ``dropout_seed`` None = eval mode (all dropouts off)."""
    drop = _Drop(dropout_seed)
    mods = list(batch.keys())
    assert mods == [m for m in spec.avail_mod if m in batch], "dict order must follow avail_mod"
    T = batch[mods[0]]["inputs"].shape[1]
    taps: Dict[str, torch.Tensor] = {}

    mask_cat = torch.cat([batch[m]["mask"] for m in mods], dim=1)            # mm.py:98-108
    attn_cat = torch.cat([batch[m]["attn_mask"] for m in mods], dim=1)
    mod_ids = torch.cat([torch.full((T,), spec.avail_mod.index(m), dtype=torch.int16) for m in mods])
    zero_ids = torch.nonzero(mask_cat[0] == 1).flatten()                     # mm.py:147 (sample 0!)

    def side(prefix_fmt: str, side_id: int):
        toks, embs = [], []
        for m in mods:
            mi = spec.avail_mod.index(m)
            x = batch[m]["inputs"]
            if x.dim() == 2:                                                  # mm.py:248-250
                x = x.unsqueeze(-1)
            t, e = embed_tokens(P, prefix_fmt.format(m), spec, mi, x, batch[m]["timestamp"], drop,
                                px.site_id(px.SITE_EMBED, mi, side_id))
            toks.append(t)
            embs.append(e)
        tok = torch.cat(toks, dim=1)
        emb = torch.cat(embs, dim=1)
        keep = torch.ones(tok.shape[1], dtype=tok.dtype)
        keep[zero_ids] = 0.0                                                  # mm.py:149 / :171
        return tok * keep[None, :, None], emb

    enc_tok, enc_emb = side("encoder_embeddings.{}.embedder", px.SIDE_ENC)
    dec_tok, dec_emb = side("decoder_embeddings.{}.embedder", px.SIDE_DEC)
    enc_ok = encoder_allowed(attn_cat)
    dec_ok = decoder_allowed(attn_cat, mod_ids, spec)

    x = enc_tok + enc_emb                                                     # mm.py:289
    if keep_taps:
        taps["x0"] = x
    for i in range(spec.n_enc_layers):                                        # encoder_embeddings.py:106-116
        pre = f"encoder.{i}"
        h = _layer_norm(P, pre + ".ln1", x)
        x = x + attention(P, pre + ".attn", spec, h, h, enc_ok, drop,
                          px.site_id(px.SITE_ATTN_PROB, i, px.SIDE_ENC),
                          px.site_id(px.SITE_ATTN_OUT, i, px.SIDE_ENC))
        x = x + mlp(P, pre + ".mlp", spec, _layer_norm(P, pre + ".ln2", x), drop,
                    px.site_id(px.SITE_MLP, i, px.SIDE_ENC))
        if keep_taps:
            taps[f"enc{i}"] = x
    x = _layer_norm(P, "encoder_norm", x)                                     # mm.py:202

    context = _linear(P, "decoder_proj_context", x) + enc_emb                 # mm.py:292
    y = dec_tok + dec_emb                                                     # mm.py:293
    if keep_taps:
        taps["context"] = context
        taps["y0"] = y
    for i in range(spec.n_dec_layers):                                        # decoder_embeddings.py:133-147
        pre = f"decoder.{i}"
        h = _layer_norm(P, pre + ".ln1", y)
        y = y + attention(P, pre + ".attn", spec, h, h, dec_ok, drop,
                          px.site_id(px.SITE_ATTN_PROB, i, px.SIDE_DEC),
                          px.site_id(px.SITE_ATTN_OUT, i, px.SIDE_DEC))
        y = y + attention(P, pre + ".cross_attn", spec, _layer_norm(P, pre + ".query_norm", y),
                          _layer_norm(P, pre + ".context_norm", context), enc_ok, drop,
                          px.site_id(px.SITE_XATTN_PROB, i, px.SIDE_DEC),
                          px.site_id(px.SITE_XATTN_OUT, i, px.SIDE_DEC))
        y = y + mlp(P, pre + ".mlp", spec, _layer_norm(P, pre + ".ln2", y), drop,
                    px.site_id(px.SITE_MLP, i, px.SIDE_DEC))
        if keep_taps:
            taps[f"dec{i}"] = y
    y = _layer_norm(P, "decoder_norm", y)                                     # mm.py:212

    mod_loss, mod_n, mod_preds, mod_targets = {}, {}, {}, {}
    for k, m in enumerate(mods):                                              # decoder_embeddings.py:95-109
        ym = y[:, k * T:(k + 1) * T, :]
        preds = _linear(P, f"decoder_embeddings.{m}.out", ym)
        tgt = batch[m]["targets"]
        if tgt.dim() == 2:
            tgt = tgt.unsqueeze(-1)
        w = batch[m]["mask"].unsqueeze(-1).expand(tgt.shape)                  # mm.py:229
        if spec.loss_kind[m] == "poisson":                                    # mm.py:80: exp(p) - t*p
            ell = torch.exp(preds) - tgt * preds
        elif spec.loss_kind[m] == "ce":
            # categorical stream (choice / block; extension -- the reference has no CE): the loss
            # module slotted into mm.py:229-231 unchanged returns the per-element field -t_k * log_softmax(p)_k, so the
            # masked sum is the token cross-entropy and the divisor is the expanded mask count like everywhere else
            ell = -(tgt * torch.log_softmax(preds, dim=-1))
        else:                                                                 # mm.py:81
            ell = (preds - tgt) ** 2
        mod_loss[m] = (ell * w).sum()                                         # mm.py:230
        mod_n[m] = w.sum()                                                    # mm.py:231
        mod_preds[m] = preds
        mod_targets[m] = tgt
    loss = sum(mod_loss.values()) / sum(mod_n.values())                       # mm.py:237
    return OracleOut(loss, mod_loss, mod_n, mod_preds, mod_targets, taps)


def forward_backward(P: Mapping[str, torch.Tensor], spec: OracleSpec, batch, dropout_seed=None,
                     dtype=torch.float32):
    """Run forward + ``loss.backward()`` (trainer/base.py:194-195) on detached copies of ``P``;
    returns (OracleOut, {name: grad}).  Parameters aliased in ``P`` (the shared ``mod_emb``,
    mm.py:84-87) are de-duplicated by storage so their gradient sums all uses."""
    uniq: Dict[int, torch.Tensor] = {}
    Q: Dict[str, torch.Tensor] = {}
    for k, v in P.items():
        key = v.data_ptr()
        if key not in uniq:
            uniq[key] = v.detach().to(dtype).clone().requires_grad_(True)
        Q[k] = uniq[key]
    b2 = {m: {k: (t.to(dtype) if torch.is_tensor(t) and t.is_floating_point() else t)
              for k, t in d.items()} for m, d in batch.items()}
    out = forward(Q, spec, b2, dropout_seed=dropout_seed)
    out.loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in Q.items()}
    return out, grads


# ---------------------------------------------------------------------------------------------
# Linear baselines of train_baseline.py
# ---------------------------------------------------------------------------------------------

def baseline_decoder(P, inputs, targets):
    """models/baseline_decoder.py:36-49: Linear(N->nb) per time bin, MSE summed / batch."""
    preds = F.linear(inputs, P["layer.weight"], P["layer.bias"])
    return ((preds - targets) ** 2).sum() / targets.shape[0], preds


def baseline_encoder(P, inputs, targets):
    """models/baseline_encoder.py:38-53: Linear(T*nb -> T*N) on the flattened trial, PoissonNLL
    (log input, no Stirling term) summed / batch."""
    B, T, N = targets.shape
    preds = F.linear(inputs.flatten(1), P["layer.weight"], P["layer.bias"]).reshape(B, T, N)
    return (torch.exp(preds) - targets * preds).sum() / B, preds
