"""influentialrs_b200 -- B200-native (sm_100a) implementation of the InfluentialRS IRN hot path.

Drop-in classes with the reference's API (model/influentialRS.py): ``InfluentialNet``, ``IRSNN``;
``ops`` holds the tensor-level operators over the C ABI in include/irs_b200.h.
"""
from . import ops  # noqa: F401
from .irn import InfluentialNet, IRSNN, PositionalEncoding  # noqa: F401

__all__ = ["ops", "InfluentialNet", "IRSNN", "PositionalEncoding"]
