"""Mint the reference entry points' command-line contract (oracle tooling; test infrastructure).

Run in the build container only:   python -m oracle.make_cli_golden
Writes tests/golden/ref_cli_defaults.json: for image_translation.py, angle_pairing.py and
distributed_image_translation.py the flags their own ``parse_args()`` accepts and the defaults it returns
(image_translation.py:21-81, angle_pairing.py:22-72, distributed_image_translation.py:48-126), plus the log-line
format string's field order (image_translation.py:394-398).  tests/test_host_logic.py holds the re-hosted parsers to it.
"""
import importlib
import json
import sys
from pathlib import Path

from .make_golden import OUT, REF, import_reference


def defaults_of(module):
    argv = sys.argv
    sys.argv = [module.__name__]
    try:
        ns = module.parse_args()
    finally:
        sys.argv = argv
    return {k: v for k, v in sorted(vars(ns).items())}


def main():
    import_reference()
    out = {}
    for name in ("image_translation", "angle_pairing", "distributed_image_translation"):
        mod = importlib.import_module(name)
        out[name] = defaults_of(mod)
    src = (REF / "image_translation.py").read_text()
    i = src.index('log_message = (f"Iter [')
    out["log_line_source"] = " ".join(l.strip() for l in src[i:i + 520].splitlines()[:5])
    OUT.mkdir(parents=True, exist_ok=True)
    (OUT / "ref_cli_defaults.json").write_text(json.dumps(out, indent=1, sort_keys=True))
    print(json.dumps(out, indent=1)[:1500])


if __name__ == "__main__":
    main()
