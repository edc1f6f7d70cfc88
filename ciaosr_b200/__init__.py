"""ciaosr_b200 -- B200-native (sm_100a) implementation of CiaoSR's implicit
attention-in-attention upsampling head behind the reference's generator /
restorer API.  See DESIGN.md for the path, the boundary and the kernels.

Importing the package does not load the CUDA library; constructing a
generator's plan or calling the head does, and fails loudly if it is missing.
"""
from .coords import make_coord, make_cell  # noqa: F401

__all__ = ["make_coord", "make_cell"]
__version__ = "0.1.0"
