"""discogan_modernized_b200 -- the DiscoGAN train step of fasion-image-generator-project/discogan_modernized
rebuilt for B200 (sm_100a): hand-written CUDA kernels behind the reference's nn.Module API.

    from discogan_modernized_b200.model import Generator, Discriminator      # drop-in for reference model.py
    from discogan_modernized_b200.train_step import DiscoGANTrainer          # the fused train step
"""
from .model import Discriminator, Generator, family_channels  # noqa: F401
from .train_step import DiscoGANTrainer, LOSS_NAMES  # noqa: F401

__all__ = ["Generator", "Discriminator", "DiscoGANTrainer", "family_channels", "LOSS_NAMES"]
