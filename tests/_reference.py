"""Helpers to import the UNMODIFIED reference (read-only, only where /root/reference exists -- i.e. in the build
container; never on the GPU box).  Test infrastructure only."""
import contextlib
import os
import sys

REF_ROOT = "/root/reference"


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "src", "multi_modal"))


@contextlib.contextmanager
def _cwd(path):
    old = os.getcwd()
    os.chdir(path)
    try:
        yield
    finally:
        os.chdir(old)


def load_config(overrides=None):
    """mm.yaml through the reference's own config_utils (train_multi_modal.py:43-48)."""
    sys.dont_write_bytecode = True
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)
    from utils.config_utils import config_from_kwargs, update_config  # noqa
    with _cwd(REF_ROOT):
        config = config_from_kwargs({"model": "include:src/configs/multi_modal/mm.yaml"})
        config = update_config("src/configs/multi_modal/trainer_mm.yaml", config)
    m = config["model"]
    for path, v in (overrides or {}).items():
        d = m
        keys = path.split(".")
        for k in keys[:-1]:
            d = d[k]
        d[keys[-1]] = v
    return config


def build_reference_model(config, n_neurons, n_behaviors, avail_mod=("ap", "behavior")):
    """train_multi_modal.py:160-189 with the reference's own classes."""
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
