"""B200-native forward/backward of the masked multi-modal encoder/decoder (drop-in for the reference's
``multi_modal.mm.MultiModal``).  See DESIGN.md."""
from .config import DotDict, default_model_config, scaled_model_config  # noqa: F401
from .masker import Masker  # noqa: F401
from .model import (DecoderEmbedding, EncoderEmbedding, MultiModal, MultiModalOutput, MultiSessionMultiModal, build_model,  # noqa: F401
                    convert)
from .optim import AdamW  # noqa: F401,E402
from . import metrics, sparse  # noqa: F401,E402
from .baselines import BaselineDecoder, BaselineEncoder  # noqa: F401,E402
from .dropin import install, installed, uninstall  # noqa: F401,E402
