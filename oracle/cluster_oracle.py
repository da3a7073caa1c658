"""CPU oracle for the reflectance clustering path.  TEST INFRASTRUCTURE ONLY (see nerf_oracle.py).

Restates object_level/cluster.py (== SSR/training/cluster.py for these functions) and the
scikit-learn routines it calls (pin in the reference: scikit-learn==0.23.2, requirements.txt:8;
installed here: see sklearn.__version__; algorithm: sklearn/cluster/_mean_shift.py
estimate_bandwidth, get_bin_seeds, _mean_shift_single_seed, MeanShift.fit).  Pinned by
tests/test_cluster_oracle.py against (a) golden vectors produced by the unmodified reference
class on the CPU and (b) the installed scikit-learn.
"""
import numpy as np
import torch


def map_color(rgb, intensity_factor=0.5):
    """Cluster.mapping_color, cluster.py:266-275: (I/3*f, g/I, b/I), I = r+g+b, unguarded."""
    rgb = torch.as_tensor(rgb, dtype=torch.float32)
    inten = torch.sum(rgb, dim=-1)
    out = torch.zeros_like(rgb)
    out[..., 0] = inten / 3.0 * intensity_factor
    out[..., 1] = rgb[..., 1] / inten
    out[..., 2] = rgb[..., 2] / inten
    return out


def inv_map_color(d, intensity_factor=0.5):
    """Cluster.inv_mapping_color, cluster.py:277-283."""
    d = torch.as_tensor(d, dtype=torch.float32)
    inten = d[..., 0] * 3.0 / intensity_factor
    g, b = d[..., 1] * inten, d[..., 2] * inten
    return torch.stack([inten - g - b, g, b], dim=-1)


def nearest_anchor(anchors, d_rgb):
    """compute_dist + argmin, cluster.py:241-252: |a|^2 + |p|^2 - 2 a.p, first minimum."""
    a, b = torch.as_tensor(anchors, dtype=torch.float32), torch.as_tensor(d_rgb, dtype=torch.float32)
    dist = torch.sum(a ** 2, dim=1).unsqueeze(1) + torch.sum(b ** 2, dim=1).unsqueeze(0) - 2 * a.mm(b.t())
    return torch.argmin(dist, dim=0).long(), dist


def choose_anchors(pixels, labels, leaf=0.01, size=100):
    """cluster.py:150-176 stated deterministically: per occupied voxel the pixel closest to the
    voxel centre (ties: lowest index), anchors in ascending voxel order."""
    pixels = torch.as_tensor(pixels, dtype=torch.float32)
    labels = torch.as_tensor(labels).long()
    vid = torch.clamp((pixels / leaf).long(), 0, size - 1)
    center = vid * leaf + leaf / 2
    dist = torch.sum((center - pixels) ** 2, dim=1)
    flat = (vid[:, 0] * size + vid[:, 1]) * size + vid[:, 2]
    order = np.lexsort((np.arange(len(flat)), np.where(np.isnan(dist.numpy()), np.inf, dist.numpy()), flat.numpy()))
    flat_sorted = flat.numpy()[order]
    first = np.ones(len(order), dtype=bool)
    first[1:] = flat_sorted[1:] != flat_sorted[:-1]
    win = order[first]
    return pixels[win], labels[win].reshape(-1, 1), flat_sorted[first]


def estimate_bandwidth(X, quantile=0.3, n_samples=5000, random_state=0):
    """sklearn.cluster.estimate_bandwidth: subsample with RandomState(0).permutation, then the
    mean distance to the int(n*quantile)-th nearest neighbour (self included)."""
    X = np.asarray(X, dtype=np.float64)
    if n_samples is not None and X.shape[0] > n_samples:
        idx = np.random.RandomState(random_state).permutation(X.shape[0])[:n_samples]
        X = X[idx]
    k = max(1, int(X.shape[0] * quantile))
    tot = 0.0
    for i in range(0, X.shape[0], 500):
        d = np.sqrt(((X[i:i + 500, None, :] - X[None, :, :]) ** 2).sum(-1))
        tot += np.partition(d, k - 1, axis=1)[:, k - 1].sum()
    return tot / X.shape[0]


def mean_shift(X, bandwidth, max_iter=300, min_bin_freq=1):
    """MeanShift(bandwidth, bin_seeding=True).fit(X) -> (centers [K,3], labels [P])."""
    X = np.asarray(X, dtype=np.float64)
    binned = np.round(X / bandwidth)
    bins, counts = np.unique(binned, axis=0, return_counts=True)
    seeds = bins[counts >= min_bin_freq]
    seeds = X if len(seeds) == len(X) else seeds * bandwidth
    stop = 1e-3 * bandwidth
    cand = []
    for s in seeds:
        mean, it = s.copy(), 0
        while True:
            within = X[np.sqrt(((X - mean) ** 2).sum(1)) <= bandwidth]
            if len(within) == 0:
                break
            old, mean = mean, within.mean(0)
            if np.linalg.norm(mean - old) <= stop or it == max_iter:
                break
            it += 1
        if len(within) > 0:
            cand.append((tuple(mean), len(within)))
    cand = sorted(set(cand), key=lambda t: (t[1], t[0]), reverse=True)
    centers = np.array([c[0] for c in cand])
    unique = np.ones(len(centers), dtype=bool)
    for i in range(len(centers)):
        if unique[i]:
            unique[np.sqrt(((centers - centers[i]) ** 2).sum(1)) <= bandwidth] = False
            unique[i] = True
    centers = centers[unique]
    d = ((X[:, None, :] - centers[None]) ** 2).sum(-1)
    return centers, d.argmin(1)


def synthetic_albedo(P, n_modes=12, seed=0):
    """Gaussian-mixture albedo pixels of SURVEY section 8d (sigma 0.03, clipped to [0.02, 1])."""
    g = torch.Generator().manual_seed(seed)
    modes = torch.rand(n_modes, 3, generator=g) * 0.8 + 0.1
    which = torch.randint(0, n_modes, (P,), generator=g)
    px = modes[which] + 0.03 * torch.randn(P, 3, generator=g)
    return px.clamp(0.02, 1.0), which
