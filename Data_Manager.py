"""Drop-in module: `from Data_Manager import HypersimDataset, ...` as with the reference's flat layout."""
import vcg_b200  # noqa: F401
from vcg_b200.Data_Manager import *  # noqa: F401,F403
from vcg_b200.Data_Manager import (DeviceLoader, DistributedShard, HypersimDataset, SatelliteMapDataset,  # noqa: F401
                                   Summer2WinterDataset, build_transforms, create_dataloaders)
