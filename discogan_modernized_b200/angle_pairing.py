"""Re-hosted ``angle_pairing.py`` (reference :181-451): reconstruction-dominated schedule (rates 0.9/0.9) and the
feature-matching loss that skips the first feature map (:55-57,111-120); the reference file itself cannot be
imported, its ``get_gan_loss`` is taken from image_translation.py as the survey prescribes."""
from ._cli import build_parser, run_training


def parse_args(argv=None):
    return build_parser("angle_pairing").parse_args(argv)


def main(argv=None):
    return run_training(parse_args(argv), "angle_pairing")


if __name__ == "__main__":
    main()
