"""Import helpers for the UNMODIFIED reference (``baseline/_ref/src`` -- vendored by ``baseline/vendor_reference.py``
-- or ``/root/reference/src`` in the build container).  Benchmark / test infrastructure only: nothing in the product
package imports this module.

The reference's model classes import with the container's packages; ``trainer.base`` additionally wants ``matplotlib``
and ``torcheval`` (absent here), which get inert stubs in ``sys.modules`` -- only plotting and the
eval-epoch R^2 helper would touch them.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CANDIDATES = (os.path.join(ROOT, "baseline", "_ref"), "/root/reference")


def ref_root():
    for c in CANDIDATES:
        if os.path.isdir(os.path.join(c, "src", "multi_modal")):
            return c
    return None


def available() -> bool:
    return ref_root() is not None


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def _stub_missing():
    """Inert stand-ins for the plotting / metric packages the trainer module imports at the top."""
    try:
        import matplotlib  # noqa: F401
    except Exception:
        m = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        plt.subplots = plt.figure = plt.close = lambda *a, **k: None
        m.pyplot = plt
        sys.modules.setdefault("matplotlib", m)
        sys.modules.setdefault("matplotlib.pyplot", plt)
    try:
        import torcheval.metrics  # noqa: F401
    except Exception:
        import torch

        class R2Score:           # torcheval.metrics.R2Score (multioutput='uniform_average'), enough for metric_utils.py
            def __init__(self):
                self.reset()

            def reset(self):
                self.p, self.t = [], []

            def to(self, device):
                return self

            def update(self, pred, target):
                self.p.append(pred.detach().double().cpu())
                self.t.append(target.detach().double().cpu())

            def compute(self):
                p, t = torch.cat(self.p), torch.cat(self.t)
                if p.dim() == 1:
                    p, t = p[:, None], t[:, None]
                ss_res = ((t - p) ** 2).sum(0)
                ss_tot = ((t - t.mean(0)) ** 2).sum(0)
                return (1.0 - ss_res / ss_tot).mean()

        te = types.ModuleType("torcheval")
        tm = types.ModuleType("torcheval.metrics")
        tm.R2Score = R2Score
        te.metrics = tm
        sys.modules.setdefault("torcheval", te)
        sys.modules.setdefault("torcheval.metrics", tm)
    try:
        import wandb  # noqa: F401
    except Exception:
        w = types.ModuleType("wandb")
        w.log = lambda *a, **k: None
        w.Image = lambda *a, **k: None
        sys.modules.setdefault("wandb", w)


def activate(trainer: bool = False) -> str:
    """Put the reference's ``src`` on ``sys.path`` (read-only use) and return the reference root."""
    root = ref_root()
    if root is None:
        raise RuntimeError("reference tree unavailable: run `python baseline/vendor_reference.py` in the build container")
    sys.dont_write_bytecode = True
    src = os.path.join(root, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    if trainer:
        _stub_missing()
    return root


def load_config(overrides=None, trainer_overrides=None):
    """mm.yaml + trainer_mm.yaml through the reference's own config_utils (train_multi_modal.py:43-48)."""
    root = activate()
    from utils.config_utils import config_from_kwargs, update_config  # noqa
    with _cwd(root):
        config = config_from_kwargs({"model": "include:src/configs/multi_modal/mm.yaml"})
        config = update_config("src/configs/multi_modal/trainer_mm.yaml", config)
    for base, upd in ((config["model"], overrides), (config, trainer_overrides)):
        for path, v in (upd or {}).items():
            d = base
            keys = path.split(".")
            for k in keys[:-1]:
                d = d[k]
            d[keys[-1]] = v
    return config


def build_reference_model(config, n_neurons, n_behaviors, avail_mod=("ap", "behavior")):
    """train_multi_modal.py:160-189 with the reference's own classes."""
    activate()
    from multi_modal.mm import MultiModal
    from multi_modal.encoder_embeddings import EncoderEmbedding
    from multi_modal.decoder_embeddings import DecoderEmbedding
    enc, dec = {}, {}
    for mod in avail_mod:
        enc[mod] = EncoderEmbedding(hidden_size=config.model.encoder.transformer.hidden_size,
                                    n_channel=n_neurons if mod == "ap" else n_behaviors, config=config.model.encoder)
    for mod in avail_mod:
        c = n_neurons if mod == "ap" else n_behaviors
        dec[mod] = DecoderEmbedding(hidden_size=config.model.decoder.transformer.hidden_size, n_channel=c,
                                    output_channel=c, config=config.model.decoder)
    return MultiModal(enc, dec, avail_mod=list(avail_mod), config=config.model, share_modality_embeddings=True)


class StubAccelerator:
    """The one attribute of ``accelerate.Accelerator`` the trainer reads (``trainer/base.py:55``)."""

    def __init__(self, device):
        import torch
        self.device = torch.device(device)


def make_trainer(model, train_loader, eval_loader, optimizer, lr_scheduler, config, device, log_dir, num_neurons,
                 mixed_training=True):
    """train_multi_modal.py:212-229 -> trainer/make.py:3-16 with the unmodified ``MultiModalTrainer``."""
    activate(trainer=True)
    from trainer.make import make_multimodal_trainer
    return make_multimodal_trainer(
        model=model, train_dataloader=train_loader, eval_dataloader=eval_loader, optimizer=optimizer,
        accelerator=StubAccelerator(device), lr_scheduler=lr_scheduler, config=config, log_dir=log_dir,
        num_neurons=num_neurons, mixed_training=mixed_training, avail_mod=list(model.avail_mod),
        modal_filter={"input": list(model.avail_mod), "output": list(model.avail_mod)}, mod_to_indx=model.mod_to_indx)
