"""`pyvb` import name for the B200-native drop-in (pyvb_b200): the reference's scripts say

    from pyvb import nodes, Network            # /root/reference/examples/PCA_missing_data.py:7

and find here the same surface (src/pyvb/__init__.py:2-3: `nodes`, `Network`), backed by the CUDA plate engine.
"""
import sys

from pyvb_b200 import LDSEngine, Network, PlateEngine, nodes, set_default_mode  # noqa: F401
from pyvb_b200 import network  # noqa: F401

sys.modules[__name__ + ".nodes"] = nodes          # `import pyvb.nodes`, `from pyvb.nodes import Gaussian`
sys.modules[__name__ + ".network"] = network
__all__ = ["nodes", "Network"]
