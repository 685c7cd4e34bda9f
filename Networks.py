"""Drop-in module: `from Networks import CycleVAEGAN, Encoder, ...` as with the reference's flat layout
(its Networks.py), served by the B200-native package."""
import vcg_b200  # noqa: F401  (registers the package)
from vcg_b200.Networks import *  # noqa: F401,F403
from vcg_b200.Networks import L, S, set_eps_source  # noqa: F401
from vcg_b200.plan import get_precision, no_wgrad, set_precision  # noqa: F401
