"""Re-hosted ``image_translation.py`` (reference :211-437): same flags, log line and checkpoints, the iteration
body (:335-390) executed by ``DiscoGANTrainer`` on the B200 kernels."""
from ._cli import build_parser, run_training


def parse_args(argv=None):
    return build_parser("image_translation").parse_args(argv)


def main(argv=None):
    return run_training(parse_args(argv), "image_translation")


if __name__ == "__main__":
    main()
