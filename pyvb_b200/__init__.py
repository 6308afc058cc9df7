"""pyvb_b200 -- B200-native VB-PCA (missing data) hot path behind pyvb's node/network API.

    from pyvb_b200 import nodes, Network          # drop-in for `from pyvb import nodes, Network`
    from pyvb_b200 import PlateEngine             # plated API for N too large for one object per row
    from pyvb_b200 import LDSEngine               # batched VB smoother for linear dynamic systems (config 5)
"""
from . import nodes
from .network import Network
from .engine import PlateEngine
from .lds import LDSEngine
from .plate import set_default_mode

__all__ = ["nodes", "Network", "PlateEngine", "LDSEngine", "set_default_mode"]
