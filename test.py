"""Drop-in CLI: `python test.py --runs_dir runs` (see the package's test.py)."""
import vcg_b200  # noqa: F401
from vcg_b200.test import *  # noqa: F401,F403
from vcg_b200.test import build_parser, evaluate_models

if __name__ == "__main__":
    evaluate_models(build_parser().parse_args())
