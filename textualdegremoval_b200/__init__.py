"""textualdegremoval_b200 -- B200 (sm_100a) implementation of the restoration-network hot path of
mrluin/TextualDegRemoval: hand-written CUDA kernels behind a C ABI (csrc/, include/tdr_sm100.h) and the
host-side mirror of the reference's arch registry (archs/)."""
from .archs import define_network  # noqa: F401
from .lib import TdrError, load as load_library  # noqa: F401

__all__ = ["define_network", "load_library", "TdrError"]
