#!/usr/bin/env python
"""Benchmark of the IntrinsicNeRF ray-marching hot path on B200.

Contract (task brief): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON line (rank 0).

The job (identical for every N: strong scaling).  One step = ``--views`` (default 24) synthetic 800x800 Blender-'chair'
views of the 100-pose orbit of BASELINE config 4 (pose_spherical(theta, -30, 4), theta in linspace(-180, 180, 101)[:-1],
every (100/views)-th pose), 64 coarse + 128 fine samples, two 8x256 intrinsic MLPs (BASELINE config 2 per view; the 100
views of config 4 scaled to 24 so that one N=1 step stays at a few seconds).  Views are sharded by image over the ranks
(reference loop: object_level/run_nerf.py:142-186 render_path); every finished frame's per-ray record (13 fp32 = 52
B/ray) is all-gathered over NCCL INSIDE the timed region, on NCCL's stream, while the rank renders its next view.

  value        rays/s of the whole job (all ranks), rays resident in HBM, CUDA events, max over ranks
  e2e          the same job through the public API with host buffers: per view pinned-host rays -> H2D ->
               object_level.render(rays=...) -> D2H of the six maps
  roofline     tensor-pipe fraction of the dominant kernel (fine-pass k_mlp_tc launch), algorithmic FLOPs
               (1 318 912 FLOP/sample) / CUDA-event time / measured bf16 peak
  cpu_baseline the UNMODIFIED reference render() (oracle/_ref staged copy or /root/reference) on the host cores, bounded
               sample (N=1 only): all-core row + the as-shipped 1-thread row + BASELINE config 1
  config5      BASELINE config 5 at N ranks: SSR training step (1024 rays sharded over the ranks, render + backward +
               gradient all-reduce + Adam)

``--impl reference`` times the unmodified reference render() on the host cores (rank 0 only; other ranks exit).
"""
import argparse
import contextlib
import json
import math
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H = W = 800
N_SAMPLES, N_IMPORTANCE = 64, 128
FLOP_PER_SAMPLE = 1318912                       # SURVEY 8d, object network, GEMMs only
FLOP_PER_RAY = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * FLOP_PER_SAMPLE
BYTES_PER_RAY = 144                             # 44 B in + 100 B out
METRIC = "rays/sec (64+128 samples)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


def measured_traffic():
    """DRAM bytes per MLP row of the dominant kernel from the tracked ncu summary (profiles/mlp_tc_traffic.json,
    written by tools/summarize_profiles.py from an `ncu --set full` capture); None when no capture is tracked."""
    p = os.path.join(ROOT, "profiles", "mlp_tc_traffic.json")
    if not os.path.exists(p):
        return None
    return json.load(open(p))


def workload_config(args, world):
    """The `config` object - identical in the product arm and the reference arm."""
    return {"workload": "blender_chair_800x800_64+128", "H": H, "W": W, "n_samples": N_SAMPLES, "n_importance": N_IMPORTANCE,
            "net": "2 x NeRF(D=8,W=256,skips=[4]) intrinsic heads", "views_per_step": args.views,
            "rays_per_step": args.views * H * W,
            "parallelism": f"views sharded by image over {world} rank(s); per-view record all-gather (NCCL) inside the timed region",
            "l2": "no explicit flush: every view streams 28 MB of rays and the per-chunk scratch is far larger than the 126 MB L2"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,power.draw,power.limit"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return None
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        if not sm:
            return None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(self.rows[0][1]), "reasons": reasons, "samples": len(sm)}
        try:                                              # board power under load next to its limit (why sw_power_cap shows up)
            pw = sorted(float(r[6]) for r in self.rows if len(r) > 7)
            out["power_w"], out["power_limit_w"] = pw[len(pw) // 2], float(self.rows[0][7])
        except (ValueError, IndexError):
            pass
        return out


def view_pose(v, n_views):
    """c2w [3,4] of view v of the job: pose_spherical(theta, -30, 4) (load_blender.py:29-34) on the 100-pose orbit."""
    idx = (v * 100) // n_views
    theta = math.radians(-180.0 + 360.0 * idx / 100.0)
    phi = math.radians(-30.0)
    ct, st, cp, sp = math.cos(theta), math.sin(theta), math.cos(phi), math.sin(phi)
    trans = torch.tensor([[1., 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 4.0], [0, 0, 0, 1]])
    rphi = torch.tensor([[1., 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]])
    rth = torch.tensor([[ct, 0, -st, 0], [0, 1., 0, 0], [st, 0, ct, 0], [0, 0, 0, 1]])
    flip = torch.tensor([[-1., 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    return (flip @ rth @ rphi @ trans)[:3, :4]


def intrinsics():
    f = 0.5 * W / math.tan(0.5 * 0.6911112070083618)
    return [[f, 0.0, 0.5 * W], [0.0, f, 0.5 * H], [0.0, 0.0, 1.0]]


# --------------------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the unmodified reference render() on the host cores
# --------------------------------------------------------------------------------------------------------------------
def reference_rows(n_rays=1024, steps=1, warmup=0, extra_rows=True):
    """The unmodified reference render() on the host cores (oracle/ref_bench.py, each row in its own interpreter with
    CUDA hidden): the all-core row on a ray sample of the 800x800 view (`steps` timed renders), plus the two
    single-thread rows of BASELINE.md section 3.  Returns (cpu_baseline dict, per-step rays/s list)."""
    from oracle import ref_bench as rb
    r = rb.row_subprocess("all_cores", n_rays, steps=steps, warmup=warmup)
    if "error" in r:
        raise RuntimeError("reference arm failed: " + r["error"])
    root = r["reference_root"]
    base = {"value": r["rays_per_s"], "unit": "rays/s", "cores": r["threads"], "kind": "reference", "host_cpus": r["host_cpus"],
            "thread_probe_rays_per_s": r["thread_probe_rays_per_s"],
            "sample": f"{r['rays']} rays (regular sub-grid) of the 800x800 view, 64+128 samples, the unmodified reference "
                      f"render(rays=...) from {os.path.relpath(root, ROOT) if root.startswith(ROOT) else root}, "
                      f"torch {r['torch']} CPU, {r['threads']} threads, CUDA hidden from the process"}
    if extra_rows:
        rows = {}
        for row, rays in (("config1", 1024), ("as_shipped", 256)):
            x = rb.row_subprocess(row, rays)
            rows[row] = {k: (round(v, 2) if isinstance(v, float) else v) for k, v in x.items() if k in ("rays_per_s", "rays", "threads", "what", "error")}
        base["rows"] = rows
    return base, r["per_step"]


def run_reference(args, rank, world):
    """--impl reference: rank 0 alone times the unmodified reference; each step = one render() of a bounded ray sample."""
    if rank != 0:
        return
    base, per_step = reference_rows(n_rays=1024, steps=args.steps, warmup=args.warmup)
    v = sum(per_step) / len(per_step)
    base["value"] = v
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * 1024 / v, "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic (random-init weights, reference seed 20220414; pose_spherical Blender camera)",
           "config": workload_config(args, world),
           "step_sample": "each step renders a bounded 1024-ray sample of the workload (cost is exactly linear in rays)",
           "cpu_baseline": base, "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# --------------------------------------------------------------------------------------------------------------------
# shared timing helpers
# --------------------------------------------------------------------------------------------------------------------
class Timer:
    def __init__(self, dev, dist):
        self.dev, self.dist = dev, dist

    def barrier(self):
        torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
            torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """Total milliseconds of `steps` calls (CUDA events per step, summed), max over ranks."""
        for _ in range(warmup):
            fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            s.record()
            fn()
            e.record()
        self.barrier()
        t = torch.tensor([sum(s.elapsed_time(e) for s, e in ev)], device=self.dev, dtype=torch.float64)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


# --------------------------------------------------------------------------------------------------------------------
# BASELINE config 5: SSR training step at N ranks
# --------------------------------------------------------------------------------------------------------------------
def make_ssr_trainer(dev, C, training):
    import intrinsicnerf_b200 as inrf
    from intrinsicnerf_b200 import ssr
    torch.manual_seed(20220414)
    mk = lambda: inrf.Semantic_NeRF(True, C, D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27,  # noqa: E731
                                    use_viewdirs=True).to(dev)
    coarse, fine = mk(), mk()

    class T(ssr.SSRRenderer):
        pass
    t = T()
    t.N_samples, t.N_importance = N_SAMPLES, N_IMPORTANCE
    t.perturb, t.raw_noise_std, t.training = (1.0, 1.0, True) if training else (0.0, 0.0, False)   # SSR_room0_config.yaml:29,34
    t.white_bkgd, t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat = False, True, C, False
    t.netchunk = t.chunk = 76800
    t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
    t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
    t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
    return t, coarse, fine


def train_step_record(dev, dist, rank, world, steps=20, warmup=5, n_rays=1024, C=28):
    """One SSR training step (SSR/training/trainer.py:851-1010 `step`): 1024 rays = 512 pixels + one neighbour each,
    perturb=1, raw_noise_std=1; render coarse+fine in training mode, photometric + semantic + intrinsic losses on the
    full batch, backward, gradient all-reduce, Adam.  At N ranks the rays are sharded (parallel.ray_shard), the rendered
    maps are all-gathered with a gradient (the losses pair ray i with ray i+N/2), and one flat all-reduce sums the
    gradients of both networks."""
    from intrinsicnerf_b200 import ops, parallel, ssr
    t, coarse, fine = make_ssr_trainer(dev, C, training=True)
    if dist is not None:
        for m in (coarse, fine):
            parallel.broadcast_weights(m, 0)
    g = torch.Generator().manual_seed(7)
    Hh, Ww = 240, 320
    pix = torch.randint(0, Hh * Ww, (n_rays // 2,), generator=g)
    nb = (pix + torch.randint(0, 3, (n_rays // 2,), generator=g) - 1).clamp(0, Hh * Ww - 1)      # a neighbour per pixel (rays.py:153-172)
    pose = torch.eye(4)
    rays_all = ssr.rays_for_batch(torch.cat([pix, nb]).to(dev), pose, Hh, Ww, Ww / 2, Ww / 2, (Ww - 1) / 2, (Hh - 1) / 2, 0.1, 10.0)
    gt = torch.rand(n_rays, 3, generator=g).to(dev)
    labels = torch.randint(0, C, (n_rays,), generator=g).to(dev)
    a, b = parallel.ray_shard(n_rays, rank, world)
    rays = rays_all[a:b].contiguous()
    nets = [coarse, fine]
    opt = torch.optim.Adam([p for m in nets for p in m.parameters()], lr=5e-4)
    ce = torch.nn.functional.cross_entropy
    keys = ("rgb_fine", "rgb_coarse", "albedo_fine", "shading_fine", "residual_fine", "albedo_coarse", "shading_coarse",
            "residual_coarse", "sem_logits_fine", "sem_logits_coarse")

    def step():
        opt.zero_grad(set_to_none=True)
        out = t.render_rays(rays)
        m = parallel.gather_maps_for_loss({k: out[k] for k in keys}, n_rays) if dist is not None else out
        loss = ((m["rgb_fine"] - gt) ** 2).mean() + ((m["rgb_coarse"] - gt) ** 2).mean() \
            + 0.04 * (ce(m["sem_logits_fine"], labels) + ce(m["sem_logits_coarse"], labels))
        for sfx in ("fine", "coarse"):
            terms = ops.intrinsic_losses(None, m["albedo_" + sfx], m["shading_" + sfx], m["residual_" + sfx], gt, labels.float(), None, "ssr")
            loss = loss + terms[1:7].sum() * 0.01
        loss.backward()
        if dist is not None:
            parallel.allreduce_gradients(nets)
        opt.step()
        return loss

    tm = Timer(dev, dist)
    l0 = ops.launch_count()
    ms = tm.timed(step, steps, warmup)
    launches = (ops.launch_count() - l0) / (steps + warmup)
    torch.cuda.synchronize()
    ops.poll_status()
    if os.environ.get("INRF_BENCH_HOSTPROF") == "1":          # development aid: where the host time of an eager step goes
        import cProfile
        import pstats
        import time
        t0 = time.perf_counter()
        for _ in range(20):
            step()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"HOSTPROF issue {1e3 * (t1 - t0) / 20:.3f} ms/step, drain after the loop {1e3 * (t2 - t1):.3f} ms", file=sys.stderr)
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(20):
            step()
        pr.disable()
        torch.cuda.synchronize()
        pstats.Stats(pr, stream=sys.stderr).sort_stats("cumulative").print_stats(45)
    # the same step replayed from a CUDA graph (one rank: no collective inside): what is left when the ~150 launches of
    # the step (ours + PyTorch's loss / Adam kernels) cost no host time
    ms_graph = None
    if dist is None:
        try:
            opt_g = torch.optim.Adam([p for m in nets for p in m.parameters()], lr=5e-4, capturable=True)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())

            def gstep():
                ops.rng_epoch_bump()                 # the in-kernel RNG seeds are constants of the graph: fresh draws per replay
                opt_g.zero_grad(set_to_none=False)
                out = t.render_rays(rays)
                loss = ((out["rgb_fine"] - gt) ** 2).mean() + ((out["rgb_coarse"] - gt) ** 2).mean() \
                    + 0.04 * (ce(out["sem_logits_fine"], labels) + ce(out["sem_logits_coarse"], labels))
                for sfx in ("fine", "coarse"):
                    terms = ops.intrinsic_losses(None, out["albedo_" + sfx], out["shading_" + sfx], out["residual_" + sfx], gt, labels.float(), None, "ssr")
                    loss = loss + terms[1:7].sum() * 0.01
                loss.backward()
                opt_g.step()
            with torch.cuda.stream(side):
                for _ in range(3):
                    gstep()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                gstep()
            ms_graph = tm.timed(graph.replay, steps, warmup) / steps
            torch.cuda.synchronize()
            ops.poll_status()
        except Exception as e:                                 # capture is an optimisation: report why it was not possible
            ms_graph = f"capture failed: {type(e).__name__}: {e}"[:200]
    flop = n_rays * (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * 2 * (692224 + 128 * C) * 3          # fwd + dX + dW
    per = ms / steps
    pk = peaks()
    n_grad = sum(p.numel() for m_ in nets for p in m_.parameters())
    return {"workload": "replica_room0_train_step_1024rays_64+128_C28", "n_gpus": world, "ms_per_step": per,
            "ms_per_step_cuda_graph": ms_graph,
            "rays_per_s": n_rays / (per * 1e-3), "rays_per_step": n_rays, "rays_per_rank": b - a,
            "algorithmic_tflops": flop / (per * 1e-3) / 1e12,
            "frac_of_sustained_bf16": flop / (per * 1e-3) / 1e12 / ((pk["bf16_sustained"] or pk["bf16_tflops"]) * world),
            "libinrf_launches_per_step": launches,
            "collectives": None if dist is None else
            f"one all_gather of the {len(keys)} rendered maps as a single record (~{n_rays * (3 * 8 + 2 * C) * 4 // 1024} KB, autograd) + one in-place all_reduce per network of "
            f"{n_grad} fp32 gradients ({n_grad * 4 / 1e6:.1f} MB)",
            "includes": "ray sampling excluded; render (training mode, in-kernel stash) + losses + backward + all-reduce + Adam"}


# --------------------------------------------------------------------------------------------------------------------
# secondary workloads (not the headline): Replica room_0 shape, eval render and training step
# --------------------------------------------------------------------------------------------------------------------
def run_room0(args, rank, local, world, dev, dist):
    """BASELINE config 3.  One step = 8 frames of 320x240 (614 400 rays) per rank through the SSR renderer; `value`
    device-resident via inrf_render_fwd, `e2e` through SSRRenderer.render_rays with host rays and D2H of rgb / depth /
    semantic logits."""
    from intrinsicnerf_b200 import ops, ssr
    C, Hh, Ww, frames = 28, 240, 320, 8
    flop_per_ray = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * 2 * (692224 + 128 * C)          # SURVEY 8d, SSR network
    t, coarse, fine = make_ssr_trainer(dev, C, training=False)
    poses = torch.eye(4).repeat(frames, 1, 1)
    for i in range(frames):
        a = math.radians(45.0 * i + 5.0 * rank)
        poses[i, 0, 0], poses[i, 0, 2], poses[i, 2, 0], poses[i, 2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    rays_dev = ssr.create_rays(frames, poses, Hh, Ww, Ww / 2, Ww / 2, (Ww - 1) / 2, (Hh - 1) / 2, 0.1, 10.0).reshape(-1, 11).contiguous()
    n_rays = rays_dev.shape[0]
    rays_host = rays_dev.cpu().pin_memory()
    pc, pf = coarse.packed(), fine.packed()
    chunks = [(i, min(i + 76800, n_rays)) for i in range(0, n_rays, 76800)]

    def step_device():
        for a_, b_ in chunks:
            ops.render_chunk(rays_dev[a_:b_], pc, pf, variant=1, n_classes=C, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE,
                             white_bkgd=False, pe_scalar_factor=10.0)

    out_host = [torch.empty(n_rays, c, pin_memory=True) for c in (3, 1, C)]

    def step_e2e():
        r = rays_host.to(dev, non_blocking=True)
        with torch.no_grad(), contextlib.redirect_stdout(sys.stderr):     # (the renderer mirrors the reference's NaN report prints)
            d = t.render_rays(r)
        for dst, k in zip(out_host, ("rgb_fine", "depth_fine", "sem_logits_fine")):
            dst.copy_(d[k].reshape(n_rays, -1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    tm = Timer(dev, dist)
    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = ops.launch_count()
        ms_dev = tm.timed(step_device, args.steps, max(3, args.warmup))
        launches = (ops.launch_count() - l0) * args.steps // (args.steps + max(3, args.warmup))
        clocks = sampler.stop() if rank == 0 else None
        e2e_steps = max(2, args.steps // 2)
        ms_e2e = tm.timed(step_e2e, e2e_steps, 1)
    torch.cuda.synchronize()
    ops.poll_status()
    if rank == 0:
        pk = peaks()
        value = n_rays * world * args.steps / (ms_dev * 1e-3)
        out = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f16 operands (RN) x f32 accumulate on tcgen05; everything else f32",
               "data": "synthetic (random-init weights, reference seed 20220414; Replica-shaped cameras)",
               "config": {"workload": "replica_room0_320x240_64+128_C28", "frames_per_step_per_gpu": frames, "rays_per_step_per_gpu": n_rays,
                          "n_samples": N_SAMPLES, "n_importance": N_IMPORTANCE, "net": "2 x Semantic_NeRF(D=8,W=256,skips=[4], C=28)",
                          "parallelism": f"frames sharded by image, dp{world}, no data-path collective",
                          "l2": "no explicit flush: per-step working set >> 126 MB L2"},
               "e2e": {"value": n_rays * world * e2e_steps / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": int(n_rays * 44),
                       "d2h_bytes_per_step": int(n_rays * (4 + C) * 4),
                       "api": "SSRRenderer.render_rays(host rays) + D2H of rgb_fine, depth_fine, sem_logits_fine"},
               "gpu_launches": launches,
               "roofline": {"bound": "tensor", "kernel": "k_mlp_tc (whole step)", "achieved": value / world * flop_per_ray / 1e12,
                            "peak": pk["bf16_sustained"] or pk["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": value / world * flop_per_ray / 1e12 / (pk["bf16_sustained"] or pk["bf16_tflops"]), "traffic": None,
                            "peak_source": pk["source"] + ", sustained bf16 (whole step)"},
               "clocks": clocks, "cpu_baseline": None}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_room0_train(args, rank, local, world, dev, dist):
    """BASELINE config 5 as its own line: metric = rays/s of the SSR training step (1024 rays per step over all ranks)."""
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    rec = train_step_record(dev, dist, rank, world, steps=max(20, args.steps), warmup=max(5, args.warmup))
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        pk = peaks()
        out = {"metric": "rays/sec (64+128 samples), training step", "value": rec["rays_per_s"], "unit": "rays/s", "n_gpus": world,
               "steps": max(20, args.steps), "warmup": max(5, args.warmup), "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
               "scaling": "strong", "vs_baseline": None, "dtype": "f16 operands x f32 accumulate (tcgen05) forward and backward; Adam in f32",
               "data": "synthetic (random-init weights, seed 20220414; Replica-shaped camera; random targets)",
               "config": {"workload": rec["workload"], "rays_per_step": rec["rays_per_step"], "n_samples": N_SAMPLES,
                          "n_importance": N_IMPORTANCE, "net": "2 x Semantic_NeRF(D=8,W=256,skips=[4], C=28)",
                          "parallelism": f"rays sharded over {world} rank(s); " + (rec["collectives"] or "no collective"),
                          "l2": "working set (stash 5.25 KB/sample x 262 144 samples = 1.4 GB per step) >> 126 MB L2"},
               "e2e": None, "gpu_launches": int(rec["libinrf_launches_per_step"] * max(20, args.steps)),
               "roofline": {"bound": "tensor", "kernel": "training step (forward + dX + dW GEMMs)", "achieved": rec["algorithmic_tflops"],
                            "peak": (pk["bf16_sustained"] or pk["bf16_tflops"]) * world, "unit": "TFLOP/s", "frac": rec["frac_of_sustained_bf16"],
                            "traffic": None, "peak_source": pk["source"] + ", sustained bf16 x ranks; 3 x forward FLOPs (SURVEY 8d)"},
               "clocks": clocks, "cpu_baseline": None, "config5": rec}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--views", type=int, default=24, help="800x800 views per step (the whole job, fixed for every N)")
    ap.add_argument("--chunk", type=int, default=160000, help="rays per inrf_render_fwd call (bounds the scratch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config5", action="store_true")
    ap.add_argument("--workload", default="chair", choices=["chair", "room0", "room0_train"],
                    help="chair: the BASELINE headline config (default). room0: BASELINE config 3 - Replica room_0 shape "
                         "(320x240 frames, Semantic_NeRF with 28 classes, near 0.1 / far 10, PE scale 10), 8 frames per step. "
                         "room0_train: BASELINE config 5 - the SSR training step")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", 0))
    local = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    import intrinsicnerf_b200 as inrf
    from intrinsicnerf_b200 import object_level as ol, ops, parallel
    ops.set_default_precision(args.precision)

    if args.workload == "room0":
        return run_room0(args, rank, local, world, dev, dist)
    if args.workload == "room0_train":
        return run_room0_train(args, rank, local, world, dev, dist)

    # random-init weights of the reference architecture, reference seed (run_nerf.py:1130); identical on every rank
    torch.manual_seed(20220414)
    mk = lambda: inrf.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True).to(dev)  # noqa: E731
    coarse, fine = mk(), mk()
    embed_fn, _ = ol.get_embedder(10, 0)
    embeddirs_fn, _ = ol.get_embedder(4, 0)
    kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(embed_fn, embeddirs_fn, 65536),
              N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0., white_bkgd=True, raw_noise_std=0.)
    V = args.views
    rounds = (V + world - 1) // world                      # views per rank (the last round may be padded with an idle rank)
    mine = parallel.image_shard(V, rank, world)            # views rank, rank+world, ...
    K = intrinsics()
    n_rays = H * W
    rays_dev = [ops.get_rays_packed(H, W, K, view_pose(v, V), 2.0, 6.0, dev) for v in mine]
    pc, pf = coarse.packed(), fine.packed()
    chunks = [(i, min(i + args.chunk, n_rays)) for i in range(0, n_rays, args.chunk)]
    rec_w = 13
    rec_local = [torch.zeros(n_rays, rec_w, device=dev) for _ in range(rounds)]
    rec_all = [torch.empty(world, n_rays, rec_w, device=dev) for _ in range(rounds)] if world > 1 else None

    def step_device():
        """The job on this rank: render my views; after each view start the all-gather of its records (NCCL stream)
        and go on rendering; wait for every gather before the step ends."""
        works = []
        for k in range(rounds):
            if k < len(rays_dev):
                for a, b in chunks:
                    o = ops.render_chunk(rays_dev[k][a:b], pc, pf, white_bkgd=True, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE)
                    rec_local[k][a:b].copy_(o["rec_fine"])
            if world > 1:
                works.append(dist.all_gather_into_tensor(rec_all[k].view(world * n_rays, rec_w), rec_local[k], async_op=True))
        for w_ in works:
            w_.wait()

    # e2e: host rays in, host maps out, through object_level.render
    rays_host = [(r[:, 0:3].contiguous().cpu().pin_memory(), r[:, 3:6].contiguous().cpu().pin_memory()) for r in rays_dev]
    out_host = [torch.empty(n_rays, c, pin_memory=True) for c in (3, 1, 1, 3, 1, 3)]

    def step_e2e():
        for o_h, d_h in rays_host:
            ro, rd = o_h.to(dev, non_blocking=True), d_h.to(dev, non_blocking=True)
            with torch.no_grad():
                res = ol.render(H, W, K, chunk=args.chunk, rays=(ro, rd), ndc=False, near=2., far=6., use_viewdirs=True, **kw)
            for dst, src in zip(out_host, res[:6]):
                dst.copy_(src.reshape(n_rays, -1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    tm = Timer(dev, dist)
    warm = max(3, args.warmup)
    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        l0 = ops.launch_count()
        ms_dev = tm.timed(step_device, args.steps, warm)
        launches = (ops.launch_count() - l0) * args.steps // (args.steps + warm)      # this rank's kernels inside the timed steps
        clocks = sampler.stop() if rank == 0 else None
        e2e_steps = max(2, args.steps // 4)
        ms_e2e = tm.timed(step_e2e, e2e_steps, 1)

        # dominant kernel alone: fine-pass launch of k_mlp_tc (or the fp32 kernel) on one chunk
        k_ms = k_flops = None
        if rays_dev:
            a, b = chunks[0]
            o = ops.render_chunk(rays_dev[0][a:b], pc, pf, white_bkgd=True, want_z=True)
            zf = o["z_fine"]
            n_k = 5
            for _ in range(2):
                ops.mlp_forward_rays(pf, 0, 0, rays_dev[0][a:b], zf)
            torch.cuda.synchronize()
            ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ks.record()
            for _ in range(n_k):
                ops.mlp_forward_rays(pf, 0, 0, rays_dev[0][a:b], zf)
            ke.record()
            torch.cuda.synchronize()
            k_ms = ks.elapsed_time(ke) / n_k
            k_flops = (b - a) * (N_SAMPLES + N_IMPORTANCE) * FLOP_PER_SAMPLE
    torch.cuda.synchronize()
    ops.poll_status()                                         # a tripped watchdog / fp16-range record would invalidate the run
    config5 = None
    if not args.no_config5:
        try:                                                  # a secondary record must not take the headline line with it
            config5 = train_step_record(dev, dist, rank, world)
        except Exception as e:
            config5 = {"workload": "replica_room0_train_step_1024rays_64+128_C28", "n_gpus": world,
                       "failed": f"{type(e).__name__}: {e}"[:300]}

    comm = None
    if world > 1:
        lt = torch.tensor([launches], device=dev, dtype=torch.float64)
        dist.all_reduce(lt)
        launches = int(lt.item())
        comm = {"collective": "all_gather_into_tensor", "backend": "nccl", "calls_per_step": rounds,
                "bytes_per_call_per_rank": n_rays * rec_w * 4, "received_bytes_per_step_per_rank": rounds * (world - 1) * n_rays * rec_w * 4,
                "overlap": "async on NCCL's stream behind the next view's render; waited for inside the timed region"}
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    job_rays = V * n_rays
    value = job_rays * args.steps / (ms_dev * 1e-3)
    e2e_val = job_rays * e2e_steps / (ms_e2e * 1e-3)
    ach_tf = k_flops / (k_ms * 1e-3) / 1e12
    kernel_name = "k_mlp_tc" if args.precision == "tc" else "k_mlp_fp32"
    tr = measured_traffic()
    rows = (chunks[0][1] - chunks[0][0]) * (N_SAMPLES + N_IMPORTANCE)
    out = {
        "metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f16 operands (RN) x f32 accumulate on tcgen05; sigma head, PE, compositing, resampling in f32" if args.precision == "tc" else "f32",
        "data": "synthetic (random-init weights, reference seed 20220414; pose_spherical Blender camera)",
        "config": workload_config(args, world),
        "impl_detail": {"chunk_rays": args.chunk, "views_per_rank": len(mine), "ms_per_view": ms_dev / args.steps / max(1, rounds)},
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": int(len(mine) * n_rays * 24), "d2h_bytes_per_step": int(len(mine) * n_rays * 48),
                "bytes_are": "per rank (rank 0)", "steps": e2e_steps,
                "api": "per view: object_level.render(rays=(o,d) copied from pinned host) + D2H of rgb,disp,acc,albedo,shading,residual"},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": kernel_name, "achieved": ach_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": ach_tf / pk["bf16_tflops"],
                     "traffic": (tr["bytes_per_row"] * rows) if tr else None, "traffic_source": tr["source"] if tr else None,
                     "peak_source": pk["source"] + ", burst bf16 (kernel timed alone)",
                     "kernel_ms": k_ms, "launch_rows": rows,
                     "whole_step_frac": value / world * FLOP_PER_RAY / 1e12 / (pk["bf16_sustained"] or pk["bf16_tflops"]),
                     "hbm_frac_algorithmic": value / world * BYTES_PER_RAY / 1e9 / pk["hbm_gbs"]},
        "clocks": clocks, "comm": comm, "config5": config5,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            out["cpu_baseline"], _ = reference_rows(n_rays=2048, steps=2, warmup=1)
        except Exception as e:                                # the staged reference is test infrastructure: say so, keep the line
            out["cpu_baseline"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
    else:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
