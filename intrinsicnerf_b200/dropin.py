"""Install the B200 path under the reference's own module names (no edits to the reference).

    import intrinsicnerf_b200.dropin as dropin
    dropin.install_object_level(run_nerf_module)   # after `import run_nerf`
    dropin.install_ssr(trainer_module, model_utils_module, rays_module, semantic_nerf_module)
"""
from . import cluster as _cluster
from . import object_level as _ol
from . import ssr as _ssr

_OBJECT_NAMES = ("Embedder", "get_embedder", "NeRF", "sample_pdf", "get_rays", "get_rays_np", "ndc_rays", "batchify",
                 "run_network", "batchify_rays", "render", "render_rays", "raw2outputs", "create_nerf",
                 "compute_intrinsic_loss", "render_path", "img2mse", "to8b")


def install_object_level(run_nerf_module, helpers_module=None):
    """Rebind the hot-path names inside an imported `run_nerf` (and `run_nerf_helpers`)."""
    for name in _OBJECT_NAMES:
        setattr(run_nerf_module, name, getattr(_ol, name))
        if helpers_module is not None and hasattr(helpers_module, name):
            setattr(helpers_module, name, getattr(_ol, name))
    run_nerf_module.Cluster, run_nerf_module.Cluster_Manager = _cluster.Cluster, _cluster.Cluster_Manager
    return run_nerf_module


def install_ssr(trainer_module=None, model_utils_module=None, rays_module=None, semantic_nerf_module=None):
    if trainer_module is not None:
        _ssr.install_into(trainer_module.SSRTrainer)
        for name in ("run_network", "raw2outputs", "sample_pdf", "batchify_rays", "get_embedder", "Semantic_NeRF", "create_rays",
                     "compute_intrinsic_loss", "Cluster_Manager"):
            if hasattr(trainer_module, name):
                setattr(trainer_module, name, getattr(_ssr, name))
    if model_utils_module is not None:
        model_utils_module.run_network, model_utils_module.raw2outputs = _ssr.run_network, _ssr.raw2outputs
    if rays_module is not None:
        rays_module.sample_pdf = _ssr.sample_pdf
        rays_module.create_rays = _ssr.create_rays
    if semantic_nerf_module is not None:
        semantic_nerf_module.Semantic_NeRF, semantic_nerf_module.get_embedder = _ssr.Semantic_NeRF, _ssr.get_embedder
        semantic_nerf_module.Embedder = _ssr.Embedder
