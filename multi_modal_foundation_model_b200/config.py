"""Configuration objects accepted by the B200 path.

The reference reads its hyper-parameters from ``src/configs/multi_modal/mm.yaml`` through a
dot-access dict (reference ``src/utils/config_utils.py:6-15``).  The B200 path accepts that very
object (duck-typed: attribute *and* item access, ``in`` test) so the user's YAML keeps working; this
module only provides (a) a dot-access dict of our own for callers that do not import the reference and
(b) the default hyper-parameters of ``mm.yaml:1-79`` restated as a Python literal so tests / bench /
smoke run without the reference tree (the GPU box has no ``/root/reference``).
"""
from __future__ import annotations

import copy
from typing import Any, Dict, Mapping, Optional


class DotDict(dict):
    """dict with attribute access; nested dicts are wrapped lazily (same contract as the
    reference's ``DictConfig``, ``config_utils.py:6-15``)."""

    def __getattr__(self, name: str) -> Any:
        try:
            value = self[name]
        except KeyError as e:  # keep AttributeError semantics for hasattr()/pickle
            raise AttributeError(name) from e
        if isinstance(value, dict) and not isinstance(value, DotDict):
            value = DotDict(value)
        return value

    def __setattr__(self, name: str, value: Any) -> None:
        self[name] = value

    def __getstate__(self):
        return dict(self)

    def __setstate__(self, state):
        self.update(state)


def _embedder(n_modality: int, max_F: int) -> Dict[str, Any]:
    # mm.yaml:27-36 / 57-66
    return dict(n_modality=n_modality, n_channels=668, max_F=max_F, mult=2, pos=True,
                act="softsign", scale=1, bias=True, dropout=0.2)


def _transformer(n_layers: int, hidden_size: int, n_heads: int, inter_size: int, use_scalenorm: bool = False) -> Dict[str, Any]:
    # mm.yaml:38-48 / 68-78
    return dict(n_layers=n_layers, hidden_size=hidden_size, use_scalenorm=use_scalenorm, n_heads=n_heads,
                attention_bias=True, act="gelu", inter_size=inter_size, mlp_bias=True, dropout=0.4,
                fixup_init=True)


def default_model_config(
    n_layers: int = 5,
    hidden_size: int = 256,
    n_heads: int = 8,
    inter_size: int = 512,
    n_modality: int = 2,
    max_F: int = 100,
    mask_ratio: float = 0.3,
    decoder_sep_mask: bool = False,
    decoder_causal_mask: bool = False,
    use_scalenorm: bool = False,
    overrides: Optional[Mapping[str, Any]] = None,
) -> DotDict:
    """Hyper-parameters of the reference's ``mm.yaml`` (defaults) as a dot-access dict.

    Field-for-field restatement of ``src/configs/multi_modal/mm.yaml:1-79``; keyword arguments cover
    the fields BASELINE.json's configs vary (layers / width / heads / modalities / time bins)."""
    cfg = dict(
        model_class="MultiModal",
        use_session=False,
        masker=dict(  # mm.yaml:5-18
            force_active=True, mode="temporal", ratio=mask_ratio, zero_ratio=1.0, random_ratio=1.0,
            expand_prob=0.0, max_timespan=1, channels=None, timesteps=None, mask_regions=["all"],
            target_regions=["all"], n_mask_regions=1, causal_zero=True,
        ),
        context=dict(forward=-1, backward=-1),  # mm.yaml:20-22
        encoder=dict(
            from_pt=None,
            embedder=_embedder(n_modality, max_F),
            transformer=_transformer(n_layers, hidden_size, n_heads, inter_size, use_scalenorm),
        ),
        decoder=dict(
            from_pt=None,
            decoder_sep_mask=decoder_sep_mask,  # mm.yaml:54-55
            decoder_causal_mask=decoder_causal_mask,
            embedder=_embedder(n_modality, max_F),
            transformer=_transformer(n_layers, hidden_size, n_heads, inter_size, use_scalenorm),
        ),
    )
    if overrides:
        cfg = _merge(cfg, overrides)
    return DotDict(copy.deepcopy(cfg))


def _merge(base: Dict[str, Any], upd: Mapping[str, Any]) -> Dict[str, Any]:
    out = dict(base)
    for k, v in upd.items():
        if isinstance(v, Mapping) and isinstance(out.get(k), dict):
            out[k] = _merge(out[k], v)
        else:
            out[k] = v
    return out


def scaled_model_config() -> DotDict:
    """BASELINE.json configs[4]: 24+24 layers, d_model 1024, 16 heads, MLP 2048, 200 bins,
    spike + 4 behaviour streams."""
    return default_model_config(n_layers=24, hidden_size=1024, n_heads=16, inter_size=2048,
                                n_modality=5, max_F=200)


def cfg_get(cfg: Any, name: str, default: Any = None) -> Any:
    """Read ``name`` from a reference ``DictConfig`` / our ``DotDict`` / a plain dict."""
    try:
        if name in cfg:
            return cfg[name]
    except TypeError:
        pass
    return getattr(cfg, name, default)
