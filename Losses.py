"""Drop-in module: `from Losses import TranslationLoss, ...` as with the reference's flat layout."""
import vcg_b200  # noqa: F401
from vcg_b200.Losses import *  # noqa: F401,F403
