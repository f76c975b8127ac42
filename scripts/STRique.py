#!/usr/bin/env python
"""Drop-in entry point with the reference's command line (scripts/STRique.py of giesselmann/STRique):
`STRique.py index ...` and `STRique.py count f5Index model repeat [--out --algn --mod_model --config --t --log_level]`.
The per-read hot path runs on the GPU (strique_b200, sm_100a); see strique_b200/cli.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from strique_b200.cli import run  # noqa: E402

if __name__ == '__main__':
    run()
