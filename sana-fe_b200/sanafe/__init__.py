"""`import sanafe` on the B200 engine: the reference's package is `from sanafecpp import *` plus pure-Python helpers
(sanafe/__init__.py:1-5); this shim re-exports the drop-in pybind11 module instead, so scripts written against
`sanafe.load_arch / load_net / Network / SpikingChip ...` run unchanged with `PYTHONPATH=sana-fe_b200`. The
reference's helper modules (layers, data, viz ...) sit above this boundary and are not part of this repository."""
from sanafe_b200.sanafecpp_b200 import *  # noqa: F401,F403
from sanafe_b200.sanafecpp_b200 import (Architecture, HardwareMappingError, Network, SpikingChip, load_arch,  # noqa: F401
                                         load_net)
