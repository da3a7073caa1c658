"""Reflectance clustering with the reference's class API, on the GPU.

  Cluster_Manager   object_level/cluster.py:11-93,  SSR/training/cluster.py:12-98
  Cluster           object_level/cluster.py:97-283, SSR/training/cluster.py:101-340

The reference runs scikit-learn's ``estimate_bandwidth`` + ``MeanShift(bin_seeding=True)`` on the
host; here the distance scans (k-th neighbour distance, flat-kernel mean-shift sweeps,
nearest-centre labelling, voxel anchor selection, nearest-anchor lookup) are CUDA kernels
(csrc/cluster.cu) and only the few-hundred-element control logic (seed binning, duplicate
removal) is host code, restating sklearn/cluster/_mean_shift.py.
"""
import json
import os

import numpy as np
import torch

from . import _lib
from .ops import _f32, _ptr, _stream, check


def _i64(t):
    return t.to(torch.int64).contiguous()


def mapping_color(rgb, intensity_factor=0.5):
    rgb = _f32(rgb, "rgb").reshape(-1, 3)
    out = torch.empty_like(rgb)
    with torch.cuda.device(rgb.device):
        check(_lib.lib().inrf_mapping_color(_ptr(rgb), rgb.shape[0], float(intensity_factor), _ptr(out), _stream()))
    return out


def nearest_anchor(points, anchors, map_color=False, intensity_factor=0.5):
    points, anchors = _f32(points, "points").reshape(-1, 3), _f32(anchors, "anchors").reshape(-1, 3)
    idx = torch.empty(points.shape[0], dtype=torch.int64, device=points.device)
    with torch.cuda.device(points.device):
        check(_lib.lib().inrf_nearest_anchor(_ptr(points), points.shape[0], _ptr(anchors), anchors.shape[0],
                                             int(map_color), float(intensity_factor), _ptr(idx), _stream()))
    return idx


def estimate_bandwidth(X, quantile=0.3, n_samples=5000, random_state=0):
    """sklearn.cluster.estimate_bandwidth: mean over a subsample of the distance to the
    int(n*quantile)-th nearest neighbour (the sample itself counts as the first)."""
    X = _f32(X, "X")
    n = X.shape[0]
    if n_samples is not None and n > n_samples:
        perm = np.random.RandomState(random_state).permutation(n)[:n_samples]
        X = X[torch.from_numpy(perm).to(X.device)].contiguous()
    k = max(1, int(X.shape[0] * quantile))
    d = torch.empty(X.shape[0], dtype=torch.float32, device=X.device)
    with torch.cuda.device(X.device):
        check(_lib.lib().inrf_kth_neighbor_dist(_ptr(X), X.shape[0], _ptr(X), X.shape[0], k, _ptr(d), _stream()))
    return float(d.double().sum().item() / X.shape[0])


def mean_shift(X, bandwidth, max_iter=300, min_bin_freq=1):
    """MeanShift(bandwidth, bin_seeding=True).fit(X) -> (cluster_centers [K,3], labels [P])."""
    X = _f32(X, "X")
    dev = X.device
    # get_bin_seeds: occupied bins of size `bandwidth`, in first-occurrence order is irrelevant
    binned = torch.round(X / bandwidth)
    bins, counts = torch.unique(binned, dim=0, return_counts=True)
    seeds = bins[counts >= min_bin_freq]
    seeds = X if seeds.shape[0] == X.shape[0] else (seeds * bandwidth).contiguous()
    Q = seeds.shape[0]
    centers = torch.empty(Q, 3, dtype=torch.float32, device=dev)
    within = torch.empty(Q, dtype=torch.int32, device=dev)
    iters = torch.empty(Q, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        check(_lib.lib().inrf_meanshift_seeds(_ptr(X), X.shape[0], _ptr(seeds), Q, float(bandwidth), int(max_iter),
                                              _ptr(centers), _ptr(within), _ptr(iters), _stream()))
    c = centers.cpu().numpy().astype(np.float64)
    w = within.cpu().numpy()
    keep = w > 0
    c, w = c[keep], w[keep]
    if c.shape[0] == 0:
        raise ValueError(f"No point was within bandwidth={bandwidth} of any seed.")
    # sort by (intensity, centre) descending; greedily drop centres within `bandwidth` of a kept one
    order = sorted(range(len(w)), key=lambda i: (w[i], tuple(c[i])), reverse=True)
    c = c[order]
    unique = np.ones(len(c), dtype=bool)
    for i in range(len(c)):
        if unique[i]:
            d = np.sqrt(((c - c[i]) ** 2).sum(1))
            unique[d <= bandwidth] = False
            unique[i] = True
    cluster_centers = torch.from_numpy(c[unique].astype(np.float32)).to(dev)
    labels = nearest_anchor(X, cluster_centers, map_color=False)
    return cluster_centers, labels


class Cluster:
    def __init__(self, device=torch.device("cuda"), intensity_factor=0.5, cluster_dir=None):
        self.batch_size = 10240       # kept for config.json compatibility; the kernels need no batching
        self.anchors = None
        self.links = None
        self.rgb_centers = None
        self.device = torch.device(device)
        self.intensity_factor = intensity_factor
        if cluster_dir is not None:
            self.load(cluster_dir)

    def load(self, cluster_dir):
        with open(os.path.join(cluster_dir, "config.json"), "r") as f:
            data = json.load(f)
        self.batch_size = data["batch_size"]
        self.intensity_factor = data["intensity_factor"]
        self.anchors = torch.Tensor(data["anchors"]).to(self.device)
        self.rgb_centers = torch.Tensor(data["rgb_centers"]).to(self.device)
        self.links = torch.Tensor(data["links"]).long().to(self.device)

    def save(self, cluster_dir):
        os.makedirs(cluster_dir, exist_ok=True)
        data = {"batch_size": self.batch_size, "intensity_factor": self.intensity_factor,
                "rgb_centers": self.rgb_centers.cpu().numpy().tolist(), "anchors": self.anchors.cpu().numpy().tolist(),
                "links": self.links.cpu().numpy().tolist()}
        path = os.path.join(cluster_dir, "config.json")
        with open(path, "w") as f:
            json.dump(data, f)
        print("successfully save cluster to:", path)
        try:                                   # colour swatches like the reference (cv2 is optional here)
            import cv2
            for i in range(self.rgb_centers.shape[0]):
                col = (255 * np.clip(np.ones((50, 50, 3)) * self.rgb_centers[i].cpu().numpy(), 0, 1)).astype(np.uint8)
                cv2.imwrite(os.path.join(cluster_dir, str(i) + ".png"), cv2.cvtColor(col, cv2.COLOR_BGR2RGB))
        except ImportError:
            pass

    # ---- colour space ---------------------------------------------------------------------
    def mapping_color(self, rgb):
        return mapping_color(rgb.to(self.device), self.intensity_factor).reshape(rgb.shape)

    def mapping_color_np(self, rgb):
        return self.mapping_color(torch.from_numpy(np.asarray(rgb, dtype=np.float32))).cpu().numpy()

    def inv_mapping_color(self, d_rgb):
        inten = d_rgb[..., 0] * 3.0 / self.intensity_factor
        g, b = d_rgb[..., 1] * inten, d_rgb[..., 2] * inten
        return torch.stack([inten - g - b, g, b], -1)

    # ---- fitting ----------------------------------------------------------------------------
    def update_center(self, pixels, quantile=0.3, n_samples=5000, band_factor=0.5):
        pixels = torch.as_tensor(np.asarray(pixels) if not torch.is_tensor(pixels) else pixels, dtype=torch.float32)
        mapped = mapping_color(pixels.to(self.device), self.intensity_factor)
        bandwidth = max(estimate_bandwidth(mapped, quantile=quantile, n_samples=n_samples) * band_factor, 0.01)
        print("bandwidth:", bandwidth)
        centers, labels = mean_shift(mapped, bandwidth)
        print("number of estimated clusters : %d" % centers.shape[0])
        self.choose_anchors(mapped, labels)
        self.rgb_centers = self.inv_mapping_color(centers).clamp(0, 1)

    def choose_anchors(self, pixels, labels):
        pixels = _f32(torch.as_tensor(pixels).to(self.device), "pixels")
        labels = _i64(torch.as_tensor(labels).to(self.device))
        P = pixels.shape[0]
        vox = torch.empty(100 ** 3, dtype=torch.int64, device=self.device)
        cap = min(P, 100 ** 3)
        anchors = torch.empty(cap, 3, dtype=torch.float32, device=self.device)
        links = torch.empty(cap, dtype=torch.int64, device=self.device)
        n = torch.zeros(1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            check(_lib.lib().inrf_choose_anchors(_ptr(pixels), _ptr(labels), P, _ptr(vox), _ptr(anchors), _ptr(links),
                                                 _ptr(n), _stream()))
        k = int(n.item())
        self.anchors = anchors[:k].clone()
        self.links = links[:k].clone().reshape(-1, 1)
        print("after merge:", self.anchors.shape)

    # ---- lookup -----------------------------------------------------------------------------
    def _dest(self, rgb, want_rgb):
        rgb = _f32(rgb.to(self.device), "rgb").reshape(-1, 3)
        P = rgb.shape[0]
        cls = torch.empty(P, dtype=torch.int64, device=self.device)
        out = torch.empty(P, 3, dtype=torch.float32, device=self.device) if want_rgb else None
        with torch.cuda.device(self.device):
            check(_lib.lib().inrf_dest_color(_ptr(rgb), P, _ptr(self.anchors), _ptr(self.links.reshape(-1).contiguous()),
                                             self.anchors.shape[0], _ptr(self.rgb_centers.contiguous()),
                                             self.rgb_centers.shape[0], float(self.intensity_factor), _ptr(out),
                                             _ptr(cls), _stream()))
        return out, cls

    def dest_color(self, rgb):
        return torch.squeeze(self._dest(rgb, True)[0])

    def dest_class(self, rgb):
        return self._dest(rgb, False)[1].reshape(-1, 1)

    def nearest_anchor(self, d_rgb):
        return nearest_anchor(d_rgb, self.anchors, map_color=False)


class Cluster_Manager:
    """Per-semantic-class list of clusters.  ``ssr_semantics=True`` reproduces the SSR fork's
    two deviations (SSR/training/cluster.py:31-32, 55-59, 75-77): cluster dirs are resolved
    relative to the config dir, and class_num==1 ignores the labels."""

    def __init__(self, class_num=0, cluster_config_file=None, ssr_semantics=False, device=torch.device("cuda")):
        self.class_num = class_num
        self.clusters = []
        self.ssr_semantics = ssr_semantics
        self.device = torch.device(device)
        if cluster_config_file is not None:
            self.load(cluster_config_file)

    def load(self, cluster_config_file):
        with open(os.path.join(cluster_config_file, "clusters.json"), "r") as f:
            data = json.load(f)
        self.class_num = data["class_num"]
        configs = data["cluster_dirs"]
        assert self.class_num == len(configs)
        self.clusters = []
        for i, cfg in enumerate(configs):
            if cfg is None:
                self.clusters.append(None)
                continue
            d = os.path.join(cluster_config_file, "c" + str(i)) if self.ssr_semantics else cfg
            self.clusters.append(Cluster(device=self.device, cluster_dir=d))
        print("load cluster num:", len(self.clusters))

    def save(self, cluster_manager_dir):
        os.makedirs(cluster_manager_dir, exist_ok=True)
        dirs = []
        for i, cl in enumerate(self.clusters):
            if cl is None:
                dirs.append(None)
                continue
            d = os.path.join(cluster_manager_dir, "c" + str(i))
            cl.save(d)
            dirs.append(d)
        path = os.path.join(cluster_manager_dir, "clusters.json")
        with open(path, "w") as f:
            json.dump({"class_num": self.class_num, "cluster_dirs": dirs}, f)
        print("successfully save cluster manager to:", path)

    def update_center(self, labels, pixels, quantile=0.3, n_samples=5000, band_factor=0.5):
        print("updating clusers...")
        on_device = torch.is_tensor(pixels) and pixels.is_cuda      # render_path hands over device tensors: no host trip
        if on_device:
            labels = torch.as_tensor(labels).to(pixels.device)
        else:
            labels = np.asarray(labels.cpu() if torch.is_tensor(labels) else labels)
            pixels = np.asarray(pixels.cpu() if torch.is_tensor(pixels) else pixels)
        self.clusters = []
        for i in range(self.class_num):
            if self.ssr_semantics and self.class_num == 1:
                cls_pixels = pixels
            else:
                cls_pixels = pixels[(labels == i).reshape(-1)]
            if len(cls_pixels) == 0:
                self.clusters.append(None)
                print("no pixels belong to class:", i)
                continue
            cl = Cluster(device=self.device)
            cl.update_center(cls_pixels, quantile=quantile, n_samples=n_samples, band_factor=band_factor)
            self.clusters.append(cl)

    def _per_class(self, rgb, label, result, fn, ignore_labels=False):
        if ignore_labels and self.clusters[0] is not None:
            return fn(self.clusters[0], rgb)
        for i in range(self.class_num):
            if self.clusters[i] is None:
                continue
            sel = torch.squeeze(label == i).reshape(-1)
            if not bool(sel.any()):
                continue
            result[sel] = fn(self.clusters[i], rgb[sel]).reshape(result[sel].shape)
        return result

    def dest_color(self, rgb, label):
        # the SSR fork ignores the labels when there is a single class - in dest_color only
        # (SSR/training/cluster.py:75-77); its dest_class still masks by label == i (:85-97)
        return self._per_class(rgb, label, rgb.clone(), lambda c, x: c.dest_color(x),
                               ignore_labels=self.ssr_semantics and self.class_num == 1)

    def dest_class(self, rgb, label):
        res = torch.zeros([rgb.shape[0], 1], dtype=torch.long, device=rgb.device)
        return self._per_class(rgb, label, res, lambda c, x: c.dest_class(x))
