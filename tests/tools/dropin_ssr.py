"""The reference's own scene-level trainer on top of intrinsicnerf_b200.dropin.

Drives the UNMODIFIED `SSRTrainer.step` (SSR/training/trainer.py:851-1065; staged copy under oracle/_ref/ on the GPU box)
- its ray sampling (`sample_data` :627, `sampling_index` rays.py:153), loss assembly, Adam update, LR decay, checkpoint
write - after `dropin.install_ssr` has rebound render_rays / volumetric_rendering / create_ssr / render_path and the
module-level names.  The trainer object is built with the reference's own `set_params()`, `set_params_replica()` and
`init_rays()`; only `prepare_data_replica` (disk + imgviz + tensorboard) is replaced by assigning the dozen attributes it
would set, from a synthetic 3-view scene.  Then `render_path(update_cluster=True)` (:1221) and steps with the cluster term.
Prints DROPIN_SSR PASS when: the losses are finite and decrease, the checkpoint written by `step` loads (strict) into the
REFERENCE's own Semantic_NeRF, render_path returns the 12-tuple with a cluster manager, and later steps use it."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

N_STEPS = int(sys.argv[1]) if len(sys.argv) > 1 else 24
from oracle import nerf_oracle as orc, refshim  # noqa: E402

sn, mu, rays_mod, tr, tu, scl = refshim.load_ssr()
RefSemanticNeRF = sn.Semantic_NeRF                 # the reference's own class, before anything is rebound
import intrinsicnerf_b200.dropin as dropin  # noqa: E402

dropin.install_ssr(tr, mu, rays_mod, sn)
C = 6                                              # valid semantic classes (void excluded)
Hh, Ww = 48, 64
with tempfile.TemporaryDirectory() as base:
    config = {
        "experiment": {"enable_semantic": True, "convention": "opencv", "endpoint_feat": False, "save_dir": base, "height": Hh,
                       "width": Ww, "dataset_type": "replica", "scene_file": base},
        "model": {"netdepth": 8, "netwidth": 256, "netdepth_fine": 8, "netwidth_fine": 256, "chunk": "1024*32", "netchunk": "1024*32"},
        "render": {"N_rays": "32*8", "N_samples": 64, "N_importance": 128, "perturb": 1, "use_viewdirs": True, "i_embed": 0,
                   "multires": 10, "multires_views": 4, "raw_noise_std": 1, "test_viz_factor": 1, "no_batching": True,
                   "depth_range": [0.1, 10.0], "white_bkgd": False},
        "train": {"lrate": 5e-4, "lrate_decay": 250e3, "N_iters": 200000, "wgt_sem": 4e-2, "w_n": 0.01, "w_f": 0.005, "w_i1": 0.1,
                  "w_i2": 0.01, "no_cluster": False, "no_semantic_tree": False, "no_intrinsic_loss": False},
        "logging": {"step_log_print": 1000, "step_log_tfb": 1000, "step_save_ckpt": 10, "step_val": 50000, "step_vis_train": 100000},
    }
    torch.manual_seed(20220414)
    np.random.seed(20220414)
    t = tr.SSRTrainer(config)                      # reference __init__: set_params + (stubbed) TFVisualizer
    t.set_params_replica()                         # reference: intrinsics, scaled sizes, exp_config.yaml
    # ---- what prepare_data_replica (trainer.py:149-280) would set, from a synthetic scene --------------------------
    n_train = 3
    yy, xx = np.meshgrid(np.arange(Hh, dtype=np.float32), np.arange(Ww, dtype=np.float32), indexing="ij")
    sem = (1 + ((xx // 11).astype(np.int64) % C))                          # vertical stripes of classes 1..C (0 = void)
    img = np.stack([0.2 + 0.1 * sem, 0.9 - 0.1 * sem, 0.3 + 0.05 * sem], -1).astype(np.float32) * (0.6 + 0.4 * yy / Hh)[..., None]
    t.ignore_label = -1
    t.num_train = t.num_test = n_train
    t.train_ids = t.test_ids = list(range(n_train))
    t.mask_ids = np.ones(n_train)
    t.num_semantic_class, t.num_valid_semantic_class = C + 1, C
    cmap = (np.arange((C + 1) * 3).reshape(C + 1, 3) * 9 % 256).astype(np.uint8)
    t.colour_map, t.valid_colour_map = torch.from_numpy(cmap).cuda(), torch.from_numpy(cmap[1:]).cuda()
    t.train_image = t.test_image = torch.from_numpy(np.stack([img] * n_train, 0)).cuda()
    t.train_depth = t.test_depth = torch.full((n_train, Hh, Ww), 2.0).cuda()
    t.train_semantic = t.test_semantic = torch.from_numpy(np.stack([sem] * n_train, 0)).cuda()
    Ts = torch.eye(4).repeat(n_train, 1, 1)
    for i in range(n_train):
        a = np.radians(20.0 * i)
        Ts[i, 0, 0], Ts[i, 0, 2], Ts[i, 2, 0], Ts[i, 2, 2] = np.cos(a), np.sin(a), -np.sin(a), np.cos(a)
    t.train_Ts = t.test_Ts = Ts
    # ---------------------------------------------------------------------------------------------------------------------
    t.create_ssr()                                 # dropin: our Semantic_NeRF modules + Adam, the reference's attribute names
    with torch.no_grad():                          # start from an opaque field (default init renders acc ~ 0.002: nothing to learn from in 24 steps)
        for net in (t.ssr_net_coarse, t.ssr_net_fine):
            net.alpha_linear.bias += 1.0
    t.init_rays()                                  # reference code; its create_rays name is bound to ours
    losses = []
    import builtins
    real_print = builtins.print

    def run_steps(first, n):
        for g in range(first, first + n):
            before = [p.detach().clone() for p in t.ssr_net_fine.parameters()][:1]
            t.step(g)
            with torch.no_grad():                  # the step's own loss is local to it: measure what it optimises on fixed rays
                t.training = False
                out = t.render_rays(t.rays[0, ::97].float())
                t.training = True
                losses.append(float(((out["rgb_fine"] - t.train_image[0].reshape(-1, 3)[::97]) ** 2).mean()))
            assert not torch.equal(before[0], next(iter(t.ssr_net_fine.parameters()))), "Adam did not move the parameters"

    run_steps(1, N_STEPS)                          # global_step 10 and 20 write checkpoints (step_save_ckpt = 10)
    torch.cuda.synchronize()
    from intrinsicnerf_b200 import ops
    ops.poll_status()
    print("eval mse after each step", [round(v, 5) for v in losses])
    ok = all(np.isfinite(losses))
    k = max(3, N_STEPS // 5)
    ok_dec = np.mean(losses[-k:]) < np.mean(losses[:k])
    ck_dir = os.path.join(base, "checkpoints")
    ckpts = sorted(os.listdir(ck_dir)) if os.path.isdir(ck_dir) else []
    ok_ckpt = False
    if ckpts:
        ck = torch.load(os.path.join(ck_dir, ckpts[-1]), map_location="cpu")
        sn.Semantic_NeRF = RefSemanticNeRF     # the reference class resolves its own name (super(Semantic_NeRF, self)) in its module
        ref_net = RefSemanticNeRF(enable_semantic=True, num_semantic_classes=C, D=8, W=256, input_ch=63, output_ch=5, skips=[4],
                                  input_ch_views=27, use_viewdirs=True)
        ref_net.load_state_dict(ck["network_fine_state_dict"], strict=True)
        x = torch.rand(256, 3, generator=torch.Generator().manual_seed(0)) * 2 - 1
        d = torch.nn.functional.normalize(torch.randn(256, 3, generator=torch.Generator().manual_seed(1)), dim=-1)
        emb = torch.cat([orc.posenc(x, 10, 10.0), orc.posenc(d, 4)], -1)
        with torch.no_grad():
            want = ref_net(emb)
        import intrinsicnerf_b200 as inrf
        ours = inrf.Semantic_NeRF(True, C, D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True)
        ours.load_state_dict(ck["network_fine_state_dict"], strict=True)
        with torch.no_grad():
            got = ours.cuda()(emb.cuda()).cpu()
        err = float((got - want).abs().max())
        print("checkpoint", ckpts[-1], "loads into the reference Semantic_NeRF; max |ours - reference| on 256 rows:", err)
        ok_ckpt = err < 2e-3 and "optimizer_state_dict" in ck
    # ---- render_path with the cluster refresh (trainer.py:1221-1443), then steps that use the cluster term ------------
    vis = os.path.join(base, "vis")
    os.makedirs(vis, exist_ok=True)
    t.training = False
    with torch.no_grad():
        res = t.render_path(t.rays_vis, save_dir=vis, update_cluster=True, b_f=0.5)
    t.training = True
    ok_path = len(res) == 12 and res[0].shape == (n_train, Hh, Ww, 3) and res[-1] is not None and np.isfinite(res[0]).all()
    t.cluster_manager = res[-1]                    # what step() does at step_vis_train (trainer.py:1066-1070)
    n_png = len([f for f in os.listdir(vis) if f.endswith(".png")])
    run_steps(N_STEPS + 1, 4)
    torch.cuda.synchronize()
    ops.poll_status()
    ok_after = all(np.isfinite(losses[-4:]))
    print("render_path ok", ok_path, "pngs", n_png, "cluster classes", res[-1].class_num if res[-1] is not None else None)
    print("DROPIN_SSR", "PASS" if (ok and ok_dec and ok_ckpt and ok_path and ok_after and n_png > 0) else
          f"FAIL finite={ok} decreased={ok_dec} ckpt={ok_ckpt} render_path={ok_path} after={ok_after} png={n_png}")
