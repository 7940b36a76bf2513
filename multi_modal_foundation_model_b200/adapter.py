"""What the step engine needs to know about a model object, read by duck typing so that BOTH class families work:

* this package's own module classes (:mod:`model`), and
* the UNMODIFIED reference classes (``multi_modal.mm.MultiModal`` and its sub-modules) once :func:`dropin.install`
  has rebound their ``forward`` -- same objects, same ``state_dict``, same pickles, other kernels underneath.

Nothing here computes anything; unsupported settings raise ``NotImplementedError`` with the config field that caused it.
"""
from __future__ import annotations

from typing import Dict

import torch.nn as nn

LOSS_KINDS = ("poisson", "mse", "ce")


def _act_name(act) -> str:
    """'softsign' | 'identity' | 'gelu' | ... from either a config string (our classes keep ``act_name``) or the
    activation module instance the reference stores (``ACT2FN[...]`` / ``nn.Identity()``,
    encoder_embeddings.py:32, mm_utils.py:46)."""
    if isinstance(act, str):
        return act
    name = type(act).__name__.lower()
    if name in ("softsign", "identity"):
        return name
    if name in ("geluactivation", "gelu"):
        # transformers' GELUActivation(use_gelu_python=False) and nn.GELU() are the exact erf form; the tanh
        # approximation ('gelu_new', 'gelu_pytorch_tanh', nn.GELU('tanh')) is a different function
        if getattr(act, "approximate", "none") != "none":
            raise NotImplementedError("tanh-approximated GELU is not built (mm.yaml:44 ships act: gelu = erf form)")
        return "gelu"
    raise NotImplementedError(f"activation {type(act).__name__} is not built in the B200 path")


def _embedder_hp(e) -> Dict[str, object]:
    act = getattr(e, "act_name", None)
    if act is None:
        act = _act_name(e.act)
    return dict(act=act, scale=float(e.scale), dropout=float(e.dropout.p), mult_ok=e.input_dim == 2 * e.n_channels,
                input_dim=int(e.input_dim), n_channels=int(e.n_channels))


def hyper_params(model, enc_groups) -> Dict[str, object]:
    """The handful of scalars the schedule depends on (mm.yaml fields), validated to be uniform
    where the kernels assume so.  ``enc_groups``: [(prefix, encoder_embeddings, decoder_embeddings)]."""
    emb = []
    for _, enc, dec in enc_groups:
        for side, md in (("encoder", enc), ("decoder", dec)):
            for name, m in md.items():
                hp = _embedder_hp(m.embedder)
                if not hp["mult_ok"]:
                    raise NotImplementedError(
                        f"{side} embedder of {name!r}: token_embed width {hp['input_dim']} != 2 * n_channels "
                        f"({hp['n_channels']}); only embedder.mult == 2 (mm.yaml:31) is built")
                emb.append((side, name, hp))
    first = emb[0][2]
    for side, name, hp in emb:
        if hp["act"] != first["act"] or hp["scale"] != first["scale"]:
            raise NotImplementedError(f"embedder act/scale differ between modalities or sides ({side}/{name}: "
                                      f"{hp['act']}, {hp['scale']} vs {first['act']}, {first['scale']})")
    enc_drop = {hp["dropout"] for side, _, hp in emb if side == "encoder"}
    dec_drop = {hp["dropout"] for side, _, hp in emb if side == "decoder"}
    if len(enc_drop) != 1 or len(dec_drop) != 1:
        raise NotImplementedError("embedder dropout must be the same for every modality of a side")
    e0, d0 = model.encoder[0], model.decoder[0]

    def mlp_act(layer):
        a = getattr(layer.mlp, "act_name", None)
        return a if a is not None else _act_name(layer.mlp.act)

    return dict(
        embed_act=first["act"], embed_scale=first["scale"],
        embed_dropout=enc_drop.pop(), dec_embed_dropout=dec_drop.pop(),
        enc_heads=int(e0.attn.n_heads), dec_heads=int(d0.attn.n_heads),
        enc_dropout=float(e0.attn.attn_dropout), dec_dropout=float(d0.attn.attn_dropout),
        enc_act=mlp_act(e0), dec_act=mlp_act(d0),
        scalenorm=not isinstance(e0.ln1, nn.LayerNorm),
    )


def loss_kinds(model) -> Dict[str, str]:
    """modality -> 'poisson' | 'mse' | 'ce'.  Our classes keep ``loss_kind``; the reference keeps loss modules
    (``loss_mod``, mm.py:79-82: PoissonNLLLoss(log_input=True) / MSELoss; losses.TokenCrossEntropy for a categorical
    stream -- an extension, the reference has none)."""
    lk = getattr(model, "loss_kind", None)
    if lk is not None:
        return dict(lk)
    out = {}
    for m, fn in model.loss_mod.items():
        if isinstance(fn, nn.PoissonNLLLoss):
            if not fn.log_input or fn.full:
                raise NotImplementedError("PoissonNLLLoss: only log_input=True, full=False (mm.py:80) is built")
            out[m] = "poisson"
        elif isinstance(fn, nn.MSELoss):
            out[m] = "mse"
        elif getattr(fn, "b200_kind", None) in LOSS_KINDS:          # losses.TokenCrossEntropy
            out[m] = fn.b200_kind
        else:
            raise NotImplementedError(f"loss {type(fn).__name__} of modality {m!r} is not built")
    return out


def mask_stream(model) -> str:
    """Stream of the token-mask sampler (see masker.py): the masker's own ``stream`` attribute when it has one (our
    class), else the drop-in's class-level default (``dropin.install(mask_stream=...)``)."""
    mk = getattr(model, "masker", None)
    s = getattr(mk, "stream", None)
    if s is None:
        s = getattr(type(mk), "b200_stream", "reference")
    return s
