"""ccdm_b200 -- B200-native (sm_100a) implementation of the CCDM denoiser hot path.

Public surface mirrors CCDM_unified: ``Unet`` (models/unet.py), ``GaussianDiffusion`` (diffusion.py),
``LabelEmbed`` (label_embedding.py), ``Trainer`` / ``EMA`` (trainer.py, ema_pytorch.py).
The arithmetic runs in libccdm_b200.so (include/ccdm_b200.h); there is no CPU or PyTorch fallback.
Training: ``loss.backward()`` runs through ccdm_b200.train (autograd nodes over the kernels);
``ccdm_b200.train_graph.GraphedTrainStep`` replays the whole optimizer step from one CUDA graph.
``sngan_generator`` (models/sngan.py) is the one-step DMD2 generator's eval-mode forward on the same conv kernels.
``VanillaUnet`` (CCDM_vanilla/.../models/unet.py) is the GroupNorm / ADM-style UNet's forward on the same conv kernels.
"""
from .unet import Unet  # noqa: F401
from .diffusion import GaussianDiffusion, ModelPrediction  # noqa: F401
from .label_embedding import LabelEmbed  # noqa: F401
from .ema import EMA  # noqa: F401
from .trainer import Trainer  # noqa: F401
from .sngan import sngan_generator  # noqa: F401
from .vanilla_unet import VanillaUnet  # noqa: F401
from .vanilla_diffusion import VanillaGaussianDiffusion  # noqa: F401
from .vanilla_trainer import VanillaTrainer  # noqa: F401

__all__ = ["Unet", "GaussianDiffusion", "ModelPrediction", "LabelEmbed", "EMA", "Trainer", "sngan_generator", "VanillaUnet",
           "VanillaGaussianDiffusion", "VanillaTrainer"]
