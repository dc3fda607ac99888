"""influentialrs_b200 -- B200-native (sm_100a) implementation of the InfluentialRS IRN hot path.

Drop-in classes with the reference's API: ``InfluentialNet`` / ``IRSNN`` (model/influentialRS.py),
``SampleNet`` (model/uRS.py), ``Evaluator`` (model/evaluator.py), and the scoring paths of the ``SAS`` and
``Caser`` baselines (model/sas.py, model/caser.py); ``ops`` holds the tensor-level operators over the C ABI
in include/irs_b200.h.
"""
from . import ops  # noqa: F401
from .irn import InfluentialNet, IRSNN, PositionalEncoding  # noqa: F401
from .urs import SampleNet  # noqa: F401
from .evaluator import Evaluator  # noqa: F401
from .baselines import SAS, Caser  # noqa: F401

__all__ = ["ops", "InfluentialNet", "IRSNN", "PositionalEncoding", "SampleNet", "Evaluator", "SAS", "Caser"]
