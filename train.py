"""Drop-in CLI: `python train.py --architecture vae_cyclegan ...` (see the package's train.py)."""
import vcg_b200  # noqa: F401
from vcg_b200.train import *  # noqa: F401,F403
from vcg_b200.train import build_parser, main

if __name__ == "__main__":
    main(build_parser().parse_args())
