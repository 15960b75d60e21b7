"""Import shims for running the UNMODIFIED reference (/root/reference) in the dev container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/generate_*.py (run once, in the dev container,
where /root/reference is mounted) to produce the known-answer fixtures under tests/golden/, and
on the GPU box -- from the copy tools/make_ref.sh stages under oracle/_ref/ -- by the reference
arm of bench.py and by tests/test_reference_fit_gpu.py, which drives the reference's own
include_batch / include_sample with hdpgpc_b200.integration enabled.  Nothing here is imported
by the product package (hdpgpc_b200).

The reference imports several packages that are absent from this image:
matplotlib, gpytorch, pyro, plotly, wfdb, torchmetrics.  They are only needed for plotting,
data download and the one-beat GP hyper-parameter fit.  `install()` registers stub modules
for them, restates `torchmetrics.audio.SignalNoiseRatio` (the published formula
10*log10((sum(target^2)+eps)/(sum((target-preds)^2)+eps)), eps = finfo(dtype).eps), wraps
scipy's `fmin_l_bfgs_b` to drop the removed `disp=` kwarg, restores `np.PINF`, and replaces
`IterativeGaussianProcess.fit_torch` (GPI.py:610-770, gpytorch ExactGP + Adam) by the pure
torch restatement in oracle/hyperfit.py ("parity unpinned": gpytorch 1.13 is not installable
here, so that sub-step cannot be checked against the real library).
"""
import importlib
import importlib.machinery
import os
import sys
import types

REFERENCE_ROOT = "/root/reference/hdpgpc"
STAGED_ROOT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "_ref", "hdpgpc")


def find_reference():
    """Where the unmodified reference can be imported from: the mounted tree in the dev container, else the copy that
    tools/make_ref.sh staged under oracle/_ref/ (git-ignored; travels to the GPU box with the gpurun snapshot)."""
    for root in (REFERENCE_ROOT, STAGED_ROOT):
        if os.path.isdir(os.path.join(root, "hdpgpc")):
            return root
    return None


class _Stub(types.ModuleType):
    """Module whose every attribute is another stub / a dummy base class."""

    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        full = self.__name__ + "." + name
        if name[:1].isupper():
            obj = type(name, (object,), {"__init__": lambda self, *a, **k: None})
        else:
            obj = _Stub(full)
            obj.__path__ = []
            obj.__spec__ = importlib.machinery.ModuleSpec(full, None, is_package=True)
            sys.modules[full] = obj
        setattr(self, name, obj)
        return obj

    def __call__(self, *a, **k):
        return None


def _stub(name):
    if name in sys.modules:
        return sys.modules[name]
    m = _Stub(name)
    m.__path__ = []
    m.__spec__ = importlib.machinery.ModuleSpec(name, None, is_package=True)
    sys.modules[name] = m
    parent, _, child = name.rpartition(".")
    if parent:
        setattr(_stub(parent), child, m)
    return m


def install(reference_root=None):
    import numpy as np
    import torch

    reference_root = reference_root or find_reference()
    if reference_root is None:
        raise ImportError("the reference package is neither mounted (/root/reference) nor staged (tools/make_ref.sh)")

    for name in ["matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.ticker",
                 "gpytorch", "gpytorch.models", "gpytorch.variational", "gpytorch.means",
                 "gpytorch.kernels", "gpytorch.constraints", "gpytorch.likelihoods",
                 "gpytorch.mlls", "gpytorch.distributions",
                 "pyro", "pyro.contrib", "pyro.contrib.gp", "pyro.distributions",
                 "plotly", "plotly.graph_objects", "plotly.subplots", "plotly.io",
                 "plotly.offline", "plotly.express", "plotly.express.colors",
                 "wfdb", "wfdb.processing",
                 "torchmetrics", "torchmetrics.audio"]:
        _stub(name)

    class SignalNoiseRatio:
        """torchmetrics==1.6.0 SignalNoiseRatio (zero_mean=False), used at GPI_HDP.py:722-743."""

        def __call__(self, preds, target):
            eps = torch.finfo(preds.dtype).eps
            noise = target - preds
            val = (torch.sum(target ** 2, dim=-1) + eps) / (torch.sum(noise ** 2, dim=-1) + eps)
            return 10 * torch.log10(val)

    sys.modules["torchmetrics.audio"].SignalNoiseRatio = SignalNoiseRatio

    if not hasattr(np, "PINF"):
        np.PINF = np.inf

    import scipy.optimize
    if not getattr(scipy.optimize.fmin_l_bfgs_b, "_hgp_wrapped", False):
        _orig = scipy.optimize.fmin_l_bfgs_b

        def fmin_l_bfgs_b(*a, **k):
            k.pop("disp", None)
            return _orig(*a, **k)

        fmin_l_bfgs_b._hgp_wrapped = True
        scipy.optimize.fmin_l_bfgs_b = fmin_l_bfgs_b

    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)

    GPI = importlib.import_module("hdpgpc.GPI")
    from oracle.hyperfit import fit_torch_restated
    GPI.IterativeGaussianProcess.fit_torch = fit_torch_restated
    hdp = importlib.import_module("hdpgpc.GPI_HDP")
    return hdp
