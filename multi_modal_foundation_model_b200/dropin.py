"""Zero-edit drop-in for the reference: ``install()`` rebinds, at CLASS level, the ``forward`` of the reference's own
``multi_modal.mm.MultiModal`` (``src/multi_modal/mm.py:242-308``) to the B200 step engine.

    import multi_modal_foundation_model_b200 as mmfm
    mmfm.install()                       # one line, before (or after) the model is built; train_multi_modal.py untouched

Everything else stays the reference's: the class objects (``multi_modal.mm.MultiModal``, ``EncoderEmbedding`` ...),
constructor, parameters, ``state_dict`` keys, ``MultiModalOutput`` dataclass, the trainer (``trainer/base.py``) and
``torch.save({"model": model})`` -- a checkpoint written with the drop-in active carries the REFERENCE class path and
unpickles in a process that has never heard of this package (``utils/eval_utils.py:62``).  The engine handle lives in
the instance ``__dict__`` under ``_b200_engine`` and is dropped by the rebound ``__getstate__``.

``uninstall()`` restores the original methods (the tests use it to run the very same model object through the
reference's PyTorch path and through the B200 kernels).  There is no fallback: with the drop-in installed a CPU-resident
model raises.
"""
from __future__ import annotations

from typing import Any, Dict, Optional

_SAVED: Dict[str, Any] = {}


def _engine_of(model):
    eng = model.__dict__.get("_b200_engine")
    if eng is None:
        from .engine import Engine
        eng = Engine(model)
        import multi_modal.mm as ref_mm
        eng.output_cls = ref_mm.MultiModalOutput
        model.__dict__["_b200_engine"] = eng
    return eng


def _forward(self, mod_dict):
    return _engine_of(self).step(mod_dict)


def _getstate(self):
    st = self.__dict__.copy()
    st.pop("_b200_engine", None)
    return st


def install(mask_stream: str = "device") -> None:
    """Rebind the reference classes (which must be importable: ``src/`` of the reference on ``sys.path``).

    ``mask_stream``: how ``token_masking`` steps (``eval_mask is None``, mm.py:266-267) draw their masks -- ``'device'``
    (default: Philox Bernoulli field sampled inside ``mmfm_mask_prep``, no host work), ``'reference'`` (bit-exact replay
    of the reference's CPU generator consumption, 0.1-0.6 s per call) or ``'fast'``; see masker.py."""
    from . import _lib, masker
    assert mask_stream in masker.STREAMS
    _lib.lib()                                   # fail loudly now if the extension is missing
    try:
        import multi_modal.mm as ref_mm
        import models.masker as ref_masker
    except ImportError as e:
        raise ImportError("install() rebinds the reference's own classes: put the reference's src/ on sys.path first "
                          f"({e})") from e
    cls = ref_mm.MultiModal
    if "forward" not in _SAVED:
        _SAVED["forward"] = cls.forward
        _SAVED["getstate"] = cls.__dict__.get("__getstate__")
    cls.forward = _forward
    cls.__getstate__ = _getstate
    cls.b200_engine = _engine_of
    ref_masker.Masker.b200_stream = mask_stream


def uninstall() -> None:
    if "forward" not in _SAVED:
        return
    import multi_modal.mm as ref_mm
    import models.masker as ref_masker
    cls = ref_mm.MultiModal
    cls.forward = _SAVED.pop("forward")
    gs = _SAVED.pop("getstate")
    if gs is None:
        del cls.__getstate__
    else:
        cls.__getstate__ = gs
    if "b200_engine" in cls.__dict__:
        del cls.b200_engine
    if "b200_stream" in ref_masker.Masker.__dict__:
        del ref_masker.Masker.b200_stream


def installed() -> bool:
    return "forward" in _SAVED
