"""pdelab_b200 — host-side binding of the B200-native PDELab operator-evaluation path.

The compute lives in dune-pdelab_b200/csrc (hand-written sm_100a CUDA behind the C ABI of
include/pdelab_b200.h).  This package only loads that library and mirrors the reference's
GridOperator call surface; it contains no CPU implementation of the path.
"""
from .abi import *  # noqa: F401,F403
from .abi import ProblemSpec  # noqa: F401
