"""deep_dantzig_b200 -- B200-native generate -> solve -> label hot path of rodrgo/deep_dantzig.

Host side mirrors the reference's import layout (``data.randomlp_dataset``, ``data.gurobi_lp``, ``ml.models.s2v``,
``ml.utils``, ``benchmark``, ``phase_transitions``) on top of the C-ABI library ``libddb200.so``
(``include/ddb200.h``).  There is no CPU path: every compute entry point raises if the CUDA library or a
B200-class device is missing.
"""
from ._lib import DdbError, abi_version, library_path  # noqa: F401

__all__ = ['DdbError', 'abi_version', 'library_path']
