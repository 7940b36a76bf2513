"""Module classes of the B200 path: same names, constructor arguments, attribute names and ``state_dict`` keys as
the reference's ``src/multi_modal`` classes, so the reference trainer / eval code (``model(mod_dict)``,
``outputs.loss.backward()``, ``torch.save({"model": model})``) runs unchanged on top of them.

The sub-modules are parameter containers: construction order and initialisers follow the reference line by line
(``nn.Linear`` / ``nn.Embedding`` / ``nn.LayerNorm`` defaults + the Fixup-style rescale of
``encoder_embeddings.py:118-129``), so a process seeded like the reference gets identical initial weights.  All
arithmetic of ``MultiModal.forward`` runs in :mod:`engine` (hand-written sm_100a kernels through the C ABI); there
is no PyTorch fallback.

Reference map: ``MultiModal`` mm.py:33-308; ``EncoderEmbedding(Layer)`` encoder_embeddings.py:19-88;
``DecoderEmbedding(Layer)`` decoder_embeddings.py:19-109; ``EncoderLayer`` encoder_embeddings.py:91-129;
``DecoderLayer`` decoder_embeddings.py:112-160; ``Attention`` / ``CrossAttention`` / ``MLP`` mm_utils.py:42-152.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional

import torch
import torch.nn as nn

from .config import cfg_get
from .masker import Masker


@dataclass
class ModelOutput:
    """models/model_output.py:11-17"""
    loss: Optional[torch.Tensor] = None
    n_examples: Optional[torch.Tensor] = None

    def to_dict(self):
        return {k: getattr(self, k) for k in self.__dataclass_fields__.keys()}


@dataclass
class MultiModalOutput(ModelOutput):
    """mm.py:24-30"""
    loss: Optional[torch.Tensor] = None
    mod_loss: Optional[Dict[str, torch.Tensor]] = None
    mod_n_examples: Optional[Dict[str, torch.Tensor]] = None
    mod_preds: Optional[Dict[str, torch.Tensor]] = None
    mod_targets: Optional[Dict[str, torch.Tensor]] = None


def _only_via_model(name: str):
    raise NotImplementedError(
        f"{name}.forward is not a stand-alone op in the B200 path: the whole encoder/decoder step is one fused "
        "kernel schedule driven by MultiModal.forward (see engine.py).")


class _EmbeddingLayer(nn.Module):
    """encoder_embeddings.py:19-42 == decoder_embeddings.py:19-41"""

    def __init__(self, hidden_size: int, n_channels: int, config):
        super().__init__()
        self.bias = cfg_get(config, "bias")
        self.n_channels = n_channels
        self.input_dim = n_channels * cfg_get(config, "mult")
        self.token_embed = nn.Linear(self.n_channels, self.input_dim, bias=self.bias)
        self.projection = nn.Linear(self.input_dim, hidden_size)
        self.act_name = cfg_get(config, "act")
        scale = cfg_get(config, "scale")
        self.scale = hidden_size ** 0.5 if scale is None else scale
        self.mod_emb = nn.Embedding(cfg_get(config, "n_modality"), hidden_size)
        self.pos = cfg_get(config, "pos")
        if self.pos:
            self.pos_embed = nn.Embedding(cfg_get(config, "max_F"), hidden_size)
        self.dropout = nn.Dropout(cfg_get(config, "dropout"))

    def forward(self, d):
        _only_via_model(type(self).__name__)


class EncoderEmbeddingLayer(_EmbeddingLayer):
    pass


class DecoderEmbeddingLayer(_EmbeddingLayer):
    pass


class EncoderEmbedding(nn.Module):
    """encoder_embeddings.py:64-88"""

    def __init__(self, n_channel: int, config, **kwargs):
        super().__init__()
        tr, emb = cfg_get(config, "transformer"), cfg_get(config, "embedder")
        self.hidden_size = cfg_get(tr, "hidden_size")
        self.n_layers = cfg_get(tr, "n_layers")
        self.max_F = cfg_get(emb, "max_F")
        self.n_channel = n_channel
        self.embedder = EncoderEmbeddingLayer(self.hidden_size, self.n_channel, emb)

    def forward(self, d):
        _only_via_model("EncoderEmbedding")


class DecoderEmbedding(nn.Module):
    """decoder_embeddings.py:65-109"""

    def __init__(self, n_channel: int, output_channel: int, config, **kwargs):
        super().__init__()
        tr, emb = cfg_get(config, "transformer"), cfg_get(config, "embedder")
        self.hidden_size = cfg_get(tr, "hidden_size")
        self.n_layers = cfg_get(tr, "n_layers")
        self.max_F = cfg_get(emb, "max_F")
        self.n_channel = n_channel
        self.output_channel = output_channel
        self.embedder = DecoderEmbeddingLayer(self.hidden_size, self.n_channel, emb)
        self.out = nn.Linear(self.hidden_size, self.output_channel)

    def forward_embed(self, d):
        _only_via_model("DecoderEmbedding")

    def out_proj(self, *a, **k):
        _only_via_model("DecoderEmbedding")


class MLP(nn.Module):
    """mm_utils.py:42-52"""

    def __init__(self, hidden_size, inter_size, act, use_bias, dropout):
        super().__init__()
        self.up_proj = nn.Linear(hidden_size, inter_size, bias=use_bias)
        self.act_name = act
        self.down_proj = nn.Linear(inter_size, hidden_size, bias=use_bias)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        _only_via_model("MLP")


class Attention(nn.Module):
    """mm_utils.py:76-114"""

    def __init__(self, idx, hidden_size, n_heads, use_bias, dropout):
        super().__init__()
        self.idx = idx
        self.hidden_size = hidden_size
        self.n_heads = n_heads
        assert self.hidden_size % self.n_heads == 0, "Hidden dim is not multiple of head size"
        self.head_size = self.hidden_size // self.n_heads
        self.query = nn.Linear(self.hidden_size, self.hidden_size, bias=use_bias)
        self.key = nn.Linear(self.hidden_size, self.hidden_size, bias=use_bias)
        self.value = nn.Linear(self.hidden_size, self.hidden_size, bias=use_bias)
        self.attn_dropout = dropout
        self.dropout = nn.Dropout(dropout)
        self.out_proj = nn.Linear(hidden_size, hidden_size, bias=use_bias)

    def forward(self, *a, **k):
        _only_via_model(type(self).__name__)


class CrossAttention(Attention):
    """mm_utils.py:118-152 (same parameters as Attention)"""


def _fixup(module: nn.Module, n_layers: int) -> None:
    """encoder_embeddings.py:118-129 / decoder_embeddings.py:149-160."""
    f = 0.67 * (n_layers) ** (-1.0 / 4.0)
    with torch.no_grad():
        for name, param in module.named_parameters():
            if name.endswith("_proj.weight"):
                param.mul_(f)
            elif name.endswith("value.weight"):
                param.copy_(f * (param * (2 ** 0.5)))


class ScaleNorm(nn.Module):
    """mm_utils.py:31-39: y = x * scale / max(||x||, eps) with ONE learned scalar (parameter container; the arithmetic
    is mmfm_scalenorm_fwd / _bwd)."""

    def __init__(self, scale, eps=1e-5):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale))
        self.eps = eps

    def forward(self, x):
        _only_via_model("ScaleNorm")


def _norm(config):
    H = cfg_get(config, "hidden_size")
    return ScaleNorm(H ** 0.5) if cfg_get(config, "use_scalenorm") else nn.LayerNorm(H)


class EncoderLayer(nn.Module):
    """encoder_embeddings.py:91-129"""

    def __init__(self, idx, config):
        super().__init__()
        self.idx = idx
        H = cfg_get(config, "hidden_size")
        self.ln1 = _norm(config)
        self.attn = Attention(idx, H, cfg_get(config, "n_heads"), cfg_get(config, "attention_bias"),
                              cfg_get(config, "dropout"))
        self.ln2 = _norm(config)
        self.mlp = MLP(H, cfg_get(config, "inter_size"), cfg_get(config, "act"), cfg_get(config, "mlp_bias"),
                       cfg_get(config, "dropout"))
        if cfg_get(config, "fixup_init"):
            _fixup(self, cfg_get(config, "n_layers"))

    def forward(self, *a, **k):
        _only_via_model("EncoderLayer")


class DecoderLayer(nn.Module):
    """decoder_embeddings.py:112-160"""

    def __init__(self, idx, config):
        super().__init__()
        self.idx = idx
        H = cfg_get(config, "hidden_size")
        nh, bias, p = cfg_get(config, "n_heads"), cfg_get(config, "attention_bias"), cfg_get(config, "dropout")
        self.ln1 = _norm(config)
        self.attn = Attention(idx, H, nh, bias, p)
        self.cross_attn = CrossAttention(idx, H, nh, bias, p)
        self.query_norm = _norm(config)
        self.context_norm = _norm(config)
        self.ln2 = _norm(config)
        self.mlp = MLP(H, cfg_get(config, "inter_size"), cfg_get(config, "act"), cfg_get(config, "mlp_bias"), p)
        if cfg_get(config, "fixup_init"):
            _fixup(self, cfg_get(config, "n_layers"))

    def forward(self, *a, **k):
        _only_via_model("DecoderLayer")


class MultiModal(nn.Module):
    """Drop-in for ``multi_modal.mm.MultiModal`` (mm.py:33-308): same constructor, ``forward(mod_dict)`` and
    ``MultiModalOutput``; the step runs on hand-written sm_100a kernels (engine.py)."""

    def __init__(self, encoder_embeddings: Dict[str, nn.Module], decoder_embeddings: Dict[str, nn.Module],
                 avail_mod: List, config, share_modality_embeddings: bool = True, **kwargs):
        super().__init__()
        enc, dec = cfg_get(config, "encoder"), cfg_get(config, "decoder")
        etr, dtr = cfg_get(enc, "transformer"), cfg_get(dec, "transformer")
        self.avail_mod = avail_mod
        self.mod_to_indx = {r: i for i, r in enumerate(self.avail_mod)}
        self.decoder_sep_mask = cfg_get(dec, "decoder_sep_mask")
        self.decoder_causal_mask = cfg_get(dec, "decoder_causal_mask")
        self.n_enc_layers = cfg_get(etr, "n_layers")
        self.n_dec_layers = cfg_get(dtr, "n_layers")
        self.hidden_size = cfg_get(etr, "hidden_size")
        self.max_F = cfg_get(cfg_get(enc, "embedder"), "max_F")
        ctx = cfg_get(config, "context")
        self.context_forward = cfg_get(ctx, "forward")
        self.context_backward = cfg_get(ctx, "backward")

        self.encoder_modalities = set(encoder_embeddings.keys())
        self.encoder_embeddings = nn.ModuleDict(encoder_embeddings)
        self.decoder_modalities = set(decoder_embeddings.keys())
        self.decoder_embeddings = nn.ModuleDict(decoder_embeddings)
        if share_modality_embeddings:
            self.share_modality_embeddings()

        mk = cfg_get(config, "masker")
        self.mask = cfg_get(mk, "force_active")
        if self.mask:
            assert cfg_get(mk, "mode") in ["temporal"], \
                "Only token-wise masking is allowed for multi-modal model for now."
            self.masker = Masker(mk)

        self.encoder = nn.ModuleList([EncoderLayer(idx, etr) for idx in range(self.n_enc_layers)])
        self.encoder_norm = nn.LayerNorm(self.hidden_size)
        self.decoder_proj_context = nn.Linear(self.hidden_size, self.hidden_size)
        self.decoder = nn.ModuleList([DecoderLayer(idx, dtr) for idx in range(self.n_dec_layers)])
        self.decoder_norm = nn.LayerNorm(self.hidden_size)
        # mm.py:79-82; extra single-channel behaviour streams default to MSE (BASELINE config 5 extension)
        # (a categorical stream -- choice / block, BASELINE.json north_star; the reference has none -- takes 'ce': one-hot
        # inputs through the ordinary embedder, K-way logits out, masked cross-entropy; pass loss_kinds={'choice': 'ce'})
        self.loss_kind = {m: ("poisson" if m == "ap" else "mse") for m in avail_mod}
        self.loss_kind.update(kwargs.get("loss_kinds") or {})

        self._engine = None

    def share_modality_embeddings(self):
        for mod in self.encoder_modalities & self.decoder_modalities:
            self.decoder_embeddings[mod].embedder.mod_emb = self.encoder_embeddings[mod].embedder.mod_emb

    # -- pickling (trainer/base.py:302-308 pickles the whole module): drop device handles ---------------------
    def __getstate__(self):
        st = self.__dict__.copy()
        st["_engine"] = None
        return st

    def engine(self):
        if self._engine is None:
            from .engine import Engine
            self._engine = Engine(self)
        return self._engine

    def forward(self, mod_dict: Dict[str, Dict[str, Any]]) -> MultiModalOutput:
        return self.engine().step(mod_dict)


class MultiSessionMultiModal(MultiModal):
    """Multi-session pre-training (BASELINE.json configs[3]; SURVEY.md section 8d config 4 -- an extension, the
    reference trains one session): ONE shared transformer, per-session ``EncoderEmbedding`` / ``DecoderEmbedding``
    modules (token_embed, projection, pos_embed, mod_emb, out head -- the "stitching" layers, sized by that
    session's neuron count) selected by the batch's ``eid`` (one session per batch, trainer/base.py:65).

    ``state_dict`` keys: the shared layers keep the reference's names; session *k*'s embedders live under
    ``session_embeddings.<key>.{encoder,decoder}_embeddings.<mod>...`` -- stripping that prefix gives exactly the
    reference's single-session keys, which is how the parity test feeds the oracle."""

    def __init__(self, session_channels: Dict[str, Dict[str, int]], avail_mod: List, config,
                 share_modality_embeddings: bool = True, **kwargs):
        first = next(iter(session_channels.values()))
        enc0 = {m: EncoderEmbedding(n_channel=first[m], config=cfg_get(config, "encoder")) for m in avail_mod}
        dec0 = {m: DecoderEmbedding(n_channel=first[m], output_channel=first[m], config=cfg_get(config, "decoder"))
                for m in avail_mod}
        super().__init__(enc0, dec0, avail_mod, config, share_modality_embeddings, **kwargs)
        # the per-session modules replace the single pair of the base class
        del self.encoder_embeddings, self.decoder_embeddings
        self._session_keys: Dict[str, str] = {}
        sess = {}
        for k, (eid, chan) in enumerate(session_channels.items()):
            key = f"s{k:03d}"
            self._session_keys[str(eid)] = key
            if k == 0:
                enc, dec = enc0, dec0
            else:
                enc = {m: EncoderEmbedding(n_channel=chan[m], config=cfg_get(config, "encoder")) for m in avail_mod}
                dec = {m: DecoderEmbedding(n_channel=chan[m], output_channel=chan[m],
                                           config=cfg_get(config, "decoder")) for m in avail_mod}
                if share_modality_embeddings:
                    for m in avail_mod:
                        dec[m].embedder.mod_emb = enc[m].embedder.mod_emb
            sess[key] = nn.ModuleDict({"encoder_embeddings": nn.ModuleDict(enc),
                                       "decoder_embeddings": nn.ModuleDict(dec)})
        self.session_embeddings = nn.ModuleDict(sess)

    def session_key(self, eid) -> Optional[str]:
        return self._session_keys.get(str(eid))

    def session_prefix(self, eid) -> str:
        return f"session_embeddings.{self._session_keys[str(eid)]}."


def build_model(n_neurons: int, n_behaviors: int, config, avail_mod=("ap", "behavior"), extra_channels=None,
                **kwargs) -> MultiModal:
    # kwargs: loss_kinds={modality: 'poisson' | 'mse' | 'ce'} overrides the reference's two defaults (mm.py:79-82)
    """Mirror of train_multi_modal.py:160-189: per-modality embedders then the model."""
    enc, dec = {}, {}
    chan = {m: (n_neurons if m == "ap" else n_behaviors) for m in avail_mod}
    if extra_channels:
        chan.update(extra_channels)
    for mod in avail_mod:
        enc[mod] = EncoderEmbedding(n_channel=chan[mod], config=cfg_get(config, "encoder"))
    for mod in avail_mod:
        dec[mod] = DecoderEmbedding(n_channel=chan[mod], output_channel=chan[mod], config=cfg_get(config, "decoder"))
    return MultiModal(enc, dec, avail_mod=list(avail_mod), config=config, share_modality_embeddings=True, **kwargs)


def convert(ref_model: nn.Module, config) -> MultiModal:
    """Build the B200 model for an existing reference ``multi_modal.mm.MultiModal`` and load its weights
    (identical ``state_dict`` keys)."""
    avail = list(ref_model.avail_mod)
    enc = {m: EncoderEmbedding(n_channel=e.n_channel, config=cfg_get(config, "encoder"))
           for m, e in ref_model.encoder_embeddings.items()}
    dec = {m: DecoderEmbedding(n_channel=e.n_channel, output_channel=e.output_channel,
                               config=cfg_get(config, "decoder"))
           for m, e in ref_model.decoder_embeddings.items()}
    model = MultiModal(enc, dec, avail_mod=avail, config=config, share_modality_embeddings=True)
    model.load_state_dict(ref_model.state_dict())
    return model
