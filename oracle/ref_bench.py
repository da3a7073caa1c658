"""Timing of the UNMODIFIED reference renderer on the host cores (TEST / BENCH INFRASTRUCTURE ONLY).

BASELINE.md section 3: what is timed is the reference's own ``render(H, W, K, chunk=32768, ..., **render_kwargs_test)``
(object_level/run_nerf.py:74-139 -> batchify_rays :59 -> render_rays :415 -> run_network :42 -> NeRF.forward,
raw2outputs :359, sample_pdf) imported from /root/reference or from the staged copy under oracle/_ref/
(oracle/build_ref.py), through oracle/refshim.py (stubs only for the I/O modules the file imports at the top).

Rows (SURVEY section 8d "CPU baseline protocol"):
  config1      32x32 full image via c2w, 64 coarse + 0 fine, 1 thread  - the PR1 reference number
  as_shipped   64+128, OMP/MKL threads = 1 (what run_nerf.py:2-3 enforces), a ray sample of the 800x800 view
  all_cores    64+128, torch.set_num_threads(n) for the best n the probe finds, same ray sample
The single-thread rows must have OMP_NUM_THREADS=1 exported BEFORE torch is imported, so they run in a
subprocess of this file (``python oracle/ref_bench.py --row as_shipped``).

Cost is exactly linear in the ray count (no early termination anywhere in render_rays), so rays/s measured on a
bounded sample of the 800x800 view is the 800x800 rays/s.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SEED = 20220414          # run_nerf.py:1130
CAMERA_ANGLE_X = 0.6911112070083618


def _reference():
    from oracle import refshim
    if not refshim.available():
        raise RuntimeError(f"reference sources not found under {refshim.REF_ROOT} (run python oracle/build_ref.py in the build container)")
    return refshim, refshim.load_object_level()


def build(n_importance):
    """The reference's create_nerf with its own seed -> render_kwargs_test (+ near/far as train() adds them)."""
    import torch
    refshim, (rn, rh, cl) = _reference()
    os.makedirs("/tmp/_inrf_ref_logs/x", exist_ok=True)
    torch.manual_seed(SEED)
    _, kw_test, *_ = rn.create_nerf(refshim.object_args(N_importance=n_importance))
    kw_test = dict(kw_test, near=2.0, far=6.0)            # run_nerf.py:705-706, 790-795
    return rn, kw_test


def camera(H, W):
    from oracle import nerf_oracle as orc
    f = 0.5 * W / math.tan(0.5 * CAMERA_ANGLE_X)          # load_blender.py:72-73
    K = [[f, 0.0, 0.5 * W], [0.0, f, 0.5 * H], [0.0, 0.0, 1.0]]   # run_nerf.py:762-767
    c2w = orc.pose_spherical(-180.0, -30.0, 4.0)[:3, :4]  # load_blender.py:29-34
    return K, c2w


def sample_rays(H, W, n):
    """n rays of the HxW view on a regular sub-grid (every image region is represented): (rays_o, rays_d) [n,3]."""
    import torch
    from oracle import nerf_oracle as orc
    K, c2w = camera(H, W)
    ro, rd = orc.get_rays(H, W, K, c2w)
    side = max(1, int(math.sqrt(n)))
    ys = torch.linspace(0, H - 1, side).long()
    xs = torch.linspace(0, W - 1, side).long()
    return K, ro[ys][:, xs].reshape(-1, 3).contiguous(), rd[ys][:, xs].reshape(-1, 3).contiguous()


def time_sample(rn, kw, H, W, n_rays, repeats=1, warmup=0):
    """Median seconds of reference render() over a ray sample of the HxW view."""
    import torch
    K, ro, rd = sample_rays(H, W, n_rays)
    ts = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            rn.render(H, W, K, chunk=32768, rays=(ro, rd), **kw)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], ro.shape[0]


def time_full_image(rn, kw, H, W, repeats=3, warmup=1):
    import torch
    K, c2w = camera(H, W)
    ts = []
    with torch.no_grad():
        for i in range(warmup + repeats):
            t0 = time.perf_counter()
            rn.render(H, W, K, chunk=32768, c2w=c2w, **kw)
            if i >= warmup:
                ts.append(time.perf_counter() - t0)
    ts.sort()
    return ts[len(ts) // 2], H * W


def probe_threads(rn, kw, H, W):
    """rays/s of a 256-ray sample for all cores, then halving (many-core hosts are often faster below the full count
    on these GEMM sizes); returns (best thread count, {threads: rays/s})."""
    import torch
    ncpu = os.cpu_count() or 1
    out = {}
    t = ncpu
    while True:
        torch.set_num_threads(t)
        time_sample(rn, kw, H, W, 64)
        dt, n = time_sample(rn, kw, H, W, 256)
        out[t] = n / dt
        if t <= 8 or t // 2 < 1:
            break
        t //= 2
    best = max(out, key=out.get)
    torch.set_num_threads(best)
    return best, out


def row_subprocess(row, n_rays, steps=1, warmup=0):
    """Every row runs in its own interpreter: CUDA is hidden (CUDA_VISIBLE_DEVICES="") so that the reference's
    `device = "cuda" if available` (run_nerf.py:27) picks the CPU on the GPU box too, and the single-thread rows get
    OMP/MKL = 1 exported before torch starts (what run_nerf.py:2-3 does when it is the entry point)."""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    if row != "all_cores":
        env.update(OMP_NUM_THREADS="1", MKL_NUM_THREADS="1")
    else:
        env.pop("OMP_NUM_THREADS", None)
        env.pop("MKL_NUM_THREADS", None)
    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--row", row, "--rays", str(n_rays), "--steps", str(steps),
                          "--warmup", str(warmup)], env=env, cwd=ROOT, capture_output=True, text=True, timeout=1800)
    for line in out.stdout.splitlines()[::-1]:
        if line.startswith("{"):
            return json.loads(line)
    return {"row": row, "error": (out.stderr or out.stdout)[-400:]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--row", required=True, choices=["config1", "as_shipped", "all_cores"])
    ap.add_argument("--rays", type=int, default=256)
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    a = ap.parse_args()
    import contextlib
    import torch
    if a.row == "all_cores":
        # the headline reference row: all host cores (best thread count by probe), `steps` renders of the ray sample
        with contextlib.redirect_stdout(sys.stderr):
            rn, kw = build(128)
            threads, table = probe_threads(rn, kw, 800, 800)
            per_step = []
            for i in range(a.warmup + a.steps):
                dt, n = time_sample(rn, kw, 800, 800, a.rays)
                if i >= a.warmup:
                    per_step.append(n / dt)
        from oracle import refshim
        print(json.dumps({"row": a.row, "rays_per_s": sum(per_step) / len(per_step), "per_step": per_step, "rays": n, "threads": threads,
                          "host_cpus": os.cpu_count(), "thread_probe_rays_per_s": {str(k): round(v, 1) for k, v in table.items()},
                          "what": f"render(800, 800, K, rays={n} of the view), 64 + 128", "torch": torch.__version__,
                          "reference_root": refshim.REF_ROOT, "cuda_visible": torch.cuda.is_available()}))
        return
    torch.set_num_threads(1)
    if a.row == "config1":
        rn, kw = build(0)
        dt, n = time_full_image(rn, kw, 32, 32, repeats=3, warmup=1)
        what = "render(32, 32, K, c2w=pose_spherical(-180,-30,4)), 64 coarse + 0 fine"
    else:
        rn, kw = build(128)
        dt, n = time_sample(rn, kw, 800, 800, a.rays, repeats=1, warmup=0)
        what = f"render(800, 800, K, rays={n} of the view), 64 + 128"
    print(json.dumps({"row": a.row, "rays_per_s": n / dt, "seconds": dt, "rays": n, "threads": torch.get_num_threads(),
                      "omp_env": os.environ.get("OMP_NUM_THREADS"), "what": what, "torch": torch.__version__}))


if __name__ == "__main__":
    main()
