"""Import shim for the upstream IntrinsicNeRF reference (test infrastructure only).

The reference lives at /root/reference in the build container; on the GPU box the nine hot-path
source files staged by oracle/build_ref.py under oracle/_ref/ (byte-identical copies, sha256 manifest)
take its place.  Used by
  * tests/golden/make_golden.py  (fixture generator, run once in the container),
  * the drop-in tests (the reference's own callers running on top of intrinsicnerf_b200.dropin),
  * bench.py --impl reference / cpu_baseline (the unmodified render() timed on the host cores).

It stubs the I/O-only imports the reference pulls in at module import time
(imageio, matplotlib, configargparse, imgviz, open3d, trimesh, skimage), makes
``torch.cuda.set_device`` a no-op and ``Tensor.cuda``/``Module.cuda`` identities on a
GPU-less host, and returns the reference modules unmodified.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
REF_ROOT = os.environ.get("INRF_REFERENCE_ROOT") or ("/root/reference" if os.path.isdir("/root/reference/object_level") else _STAGED)


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "object_level"))


def _stub(name, **attrs):
    if name in sys.modules:
        return sys.modules[name]
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__path__ = []  # behave like a package so "import a.b" works

    class _Anything:  # any attribute of a stubbed I/O module: callable, subclassable, inert
        def __init__(self, *a, **k):
            pass

        def __call__(self, *a, **k):
            return None

        def __getattr__(self, n):
            return _Anything()
    m.__getattr__ = lambda n: _Anything if not n.startswith("__") else (_ for _ in ()).throw(AttributeError(n))
    sys.modules[name] = m
    return m


def _install_stubs():
    import torch
    for name in ("imageio", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
                 "configargparse", "imgviz", "imgviz.draw", "open3d", "trimesh",
                 "skimage", "skimage.io", "skimage.measure", "skimage.transform",
                 "tensorboardX", "torch.utils.tensorboard"):
        try:
            __import__(name)
        except Exception:
            _stub(name)
    if "torch.utils.tensorboard" in sys.modules and not hasattr(sys.modules["torch.utils.tensorboard"], "SummaryWriter"):
        sys.modules["torch.utils.tensorboard"].SummaryWriter = object
    if not torch.cuda.is_available():
        torch.cuda.set_device = lambda *a, **k: None
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self


def load_object_level():
    """Returns (run_nerf, run_nerf_helpers, cluster) reference modules."""
    _install_stubs()
    p = os.path.join(REF_ROOT, "object_level")
    if p not in sys.path:
        sys.path.insert(0, p)
    keep_threads = (os.environ.get("OMP_NUM_THREADS"), os.environ.get("MKL_NUM_THREADS"))
    import run_nerf  # noqa: E402  (sets OMP/MKL=1 in os.environ; harmless after torch import)
    import run_nerf_helpers
    import cluster
    for k, v in zip(("OMP_NUM_THREADS", "MKL_NUM_THREADS"), keep_threads):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return run_nerf, run_nerf_helpers, cluster


def load_ssr():
    """Returns (semantic_nerf, model_utils, rays, trainer, training_utils, ssr_cluster)."""
    _install_stubs()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import torch
    from SSR.models import semantic_nerf, model_utils, rays
    from SSR.training import trainer, training_utils
    from SSR.training import cluster as ssr_cluster
    torch.autograd.set_detect_anomaly(False)  # reference switches it on at import
    return semantic_nerf, model_utils, rays, trainer, training_utils, ssr_cluster


def object_args(N_importance=128, netchunk=65536):
    return types.SimpleNamespace(
        multires=10, multires_views=4, i_embed=0, use_viewdirs=True,
        netdepth=8, netwidth=256, netdepth_fine=8, netwidth_fine=256,
        netchunk=netchunk, N_samples=64, N_importance=N_importance, perturb=0.,
        raw_noise_std=0., white_bkgd=True, no_reload=True, ft_path=None,
        basedir="/tmp/_inrf_ref_logs", expname="x", lrate=5e-4,
        dataset_type="blender", no_ndc=False, lindisp=False)
