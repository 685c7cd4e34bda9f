"""Drop-in module: `from utils import save_checkpoint, load_checkpoint, ...` as with the reference's flat layout."""
import vcg_b200  # noqa: F401
from vcg_b200.utils import *  # noqa: F401,F403
from vcg_b200.utils import (load_checkpoint, load_pretrained_doubleae_to_cycleae, load_pretrained_doublevae_to_cyclevae,  # noqa: F401
                            save_checkpoint, truncate_tensorboard_events)
