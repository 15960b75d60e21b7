"""hdpgpc_b200 -- B200-native (sm_100a) implementation of HDP-GPC's variational E-step hot path.

Drop-in for the E-step seam of AdrianPerezHerrero/HDP-GPC (see DESIGN.md / INTEGRATION.md):
`GPI_model` (emission scoring) and `GPI_HDP` (lead weights, HMM smoothing, hard responsibilities,
sufficient statistics, cluster_new_batch) keep the reference's method names; `EStepEngine` is the
struct-of-arrays sweep that bench.py measures.  All arithmetic runs in hand-written CUDA kernels
behind the C ABI in include/hdpgpc_b200.h; there is no CPU fallback.
"""
from ._lib import HgpError, load as load_library  # noqa: F401
from .build import build  # noqa: F401
from .model import GPI_model, LinAlgError, full_pass_weighted_batch  # noqa: F401
from .hdp import GPI_HDP, EStepEngine, LeadTables  # noqa: F401

__version__ = "0.1.0"
