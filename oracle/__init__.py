"""CPU/PyTorch restatement of the reference DiscoGAN train step.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product
path: it may be imported only by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` -- as the checker or
the CPU baseline, never as the thing measured or shipped.

Parity pinning: the reference (fasion-image-generator-project/discogan_modernized)
ships no tests, golden vectors or KATs, so the oracle is pinned
against outputs of the reference's own Python run in the build container:
``oracle/make_golden.py`` imports ``/root/reference/model.py`` and
``/root/reference/image_translation.py`` (losses), runs them on seeded inputs and
commits the vectors under ``tests/golden/``; ``tests/test_oracle.py`` replays the
oracle against those vectors.  The arithmetic itself lives in the third-party
dependency ``torch`` (un-pinned by the reference, ``requirements.txt:9-12``; docs
say 2.1.0; the oracle uses the installed 2.11.0).
"""
from .family import Generator, Discriminator, family_channels  # noqa: F401
from .losses import get_gan_loss, get_fm_loss, get_fm_loss_angle  # noqa: F401
from .step import OracleStep, synthetic_batch, build_nets  # noqa: F401
