"""The reference's own object-level training loop on top of intrinsicnerf_b200.dropin.

Runs the UNMODIFIED `run_nerf.train()` (object_level/run_nerf.py:664-1125; staged copy under oracle/_ref/ on the GPU box)
exactly as its __main__ does - CUDA default tensor type, seeds 20220414 - after `dropin.install_object_level` has rebound
the hot-path names.  Only the things outside the hot path are substituted, and none of them by editing the reference:
  * `configargparse` (not installed here) -> a 5-line argparse subclass that accepts `is_config_file`;
  * `load_blender_data` -> a synthetic 6-view scene (a shaded disk on white, alpha = object mask);
  * `trange` -> a bounded range (the loop is hard-coded to 200 001 iterations, run_nerf.py:853);
  * `tqdm.write` -> records the "[TRAIN] Iter .. Loss .." lines.
dataset_type=blender_intrinsic because of reference quirk A3 (target_m is only defined there).
Exercised through the reference's code: render(rays=..., retraw=True) (:942), compute_intrinsic_loss, loss.backward(),
Adam, checkpoint save (:1035), render_path(update_cluster=True) + cluster.save (:1062-1075), cluster.dest_color in the loop
(:947).  Prints DROPIN_OBJECT PASS when: every loss is finite, the loss decreased, the checkpoint written by the loop
loads (strict) into the REFERENCE's own NeRF class and reproduces our module's output, the cluster state was saved."""
import argparse
import os
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

N_ITERS = int(sys.argv[1]) if len(sys.argv) > 1 else 24


class _ConfigArgParser(argparse.ArgumentParser):
    def add_argument(self, *a, **k):
        k.pop("is_config_file", None)
        return super().add_argument(*a, **k)


cfg = types.ModuleType("configargparse")
cfg.ArgumentParser = _ConfigArgParser
sys.modules["configargparse"] = cfg

from oracle import nerf_oracle as orc, refshim  # noqa: E402

rn, rh, cl = refshim.load_object_level()
RefNeRF = rh.NeRF                                   # the reference's own module class, before anything is rebound
import intrinsicnerf_b200.dropin as dropin  # noqa: E402

dropin.install_object_level(rn, rh)


def synthetic_blender(datadir, half_res, testskip):
    Hs = Ws = 40
    n = 6
    focal = 0.5 * Ws / np.tan(0.5 * 0.6911112070083618)
    yy, xx = np.meshgrid(np.arange(Hs, dtype=np.float32), np.arange(Ws, dtype=np.float32), indexing="ij")
    r = np.sqrt((yy - Hs / 2) ** 2 + (xx - Ws / 2) ** 2)
    mask = (r < 14).astype(np.float32)
    imgs = np.zeros((n, Hs, Ws, 4), np.float32)
    poses = np.zeros((n, 4, 4), np.float32)
    for i in range(n):
        shade = 0.35 + 0.5 * (xx / Ws) * np.cos(i) ** 2 + 0.1 * (yy / Hs)
        imgs[i, ..., 0], imgs[i, ..., 1], imgs[i, ..., 2] = 0.8 * shade, 0.4 * shade, 0.3 * shade
        imgs[i, ..., 3] = mask
        poses[i] = np.asarray(orc.pose_spherical(-180.0 + 60.0 * i, -30.0, 4.0), np.float32)
    render_poses = torch.stack([torch.as_tensor(np.asarray(orc.pose_spherical(a, -30.0, 4.0), np.float32)) for a in (-180.0, 0.0)], 0)
    return imgs, poses, render_poses, [Hs, Ws, focal], [np.array([0, 1, 2, 3]), np.array([4]), np.array([4, 5])]


class _Tqdm:
    lines = []

    @classmethod
    def write(cls, s):
        cls.lines.append(s)


with tempfile.TemporaryDirectory() as base:
    rn.load_blender_data = synthetic_blender
    rn.trange = lambda a, b: range(a, a + N_ITERS)
    rn.tqdm = _Tqdm
    sys.argv = ["run_nerf.py", "--expname", "dropin", "--basedir", base, "--datadir", "synthetic", "--dataset_type", "blender_intrinsic",
                "--no_batching", "--use_viewdirs", "--white_bkgd", "--no_reload", "--N_samples", "64", "--N_importance", "128",
                "--N_rand", "512", "--i_print", "1", "--i_weights", "10", "--i_testset", "12", "--chunk", "32768", "--netchunk", "65536"]
    torch.set_default_tensor_type("torch.cuda.FloatTensor")     # run_nerf.py:1129-1131
    torch.manual_seed(20220414)
    np.random.seed(20220414)
    try:
        rn.train()
    finally:
        torch.set_default_tensor_type("torch.FloatTensor")
    torch.cuda.synchronize()
    from intrinsicnerf_b200 import ops
    ops.poll_status()
    losses = [float(s.split("Loss:")[1].split()[0]) for s in _Tqdm.lines if "Loss:" in s]
    print("losses", [round(v, 4) for v in losses])
    ok = len(losses) == N_ITERS and all(np.isfinite(losses))
    k = max(3, N_ITERS // 5)
    ok_dec = np.mean(losses[-k:]) < np.mean(losses[:k])
    exp = os.path.join(base, "dropin")
    ckpts = sorted(f for f in os.listdir(exp) if f.endswith(".tar"))
    ok_ckpt = False
    if ckpts:
        ck = torch.load(os.path.join(exp, ckpts[-1]), map_location="cpu")
        rh.NeRF = RefNeRF          # the reference class resolves its own name (super(NeRF, self)) in run_nerf_helpers' globals
        ref_net = RefNeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
        ref_net.load_state_dict(ck["network_fine_state_dict"], strict=True)
        import intrinsicnerf_b200 as inrf
        ours = inrf.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
        ours.load_state_dict(ck["network_fine_state_dict"], strict=True)
        ours = ours.cuda()
        x = torch.rand(256, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
        d = torch.nn.functional.normalize(torch.randn(256, 3, generator=torch.Generator().manual_seed(1)), dim=-1)
        emb = torch.cat([orc.posenc(x, 10), orc.posenc(d, 4)], -1)
        with torch.no_grad():
            want = ref_net(emb)
            got = ours(emb.cuda()).cpu()
        err = float((got - want).abs().max())
        print("checkpoint", ckpts[-1], "loads into the reference NeRF; max |ours - reference| on 256 rows:", err)
        ok_ckpt = "optimizer_state_dict" in ck and err < 2e-3
    cdir = os.path.join(exp, "cluster_{:06d}".format(12))
    ok_cluster = os.path.exists(os.path.join(cdir, "clusters.json"))
    pngs = [f for f in os.listdir(os.path.join(exp, "testset_{:06d}".format(12))) if f.endswith(".png")]
    print("testset PNGs", len(pngs), "cluster saved", ok_cluster, "ckpts", ckpts)
    ok_png = len(pngs) == 2 * 7                       # 2 test views x (rgb, a, s, res, acc, c, edit)
    print("DROPIN_OBJECT", "PASS" if (ok and ok_dec and ok_ckpt and ok_cluster and ok_png) else
          f"FAIL finite={ok} decreased={ok_dec} ckpt={ok_ckpt} cluster={ok_cluster} png={ok_png}")
