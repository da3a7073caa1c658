#!/usr/bin/env python
"""Benchmark of the IntrinsicNeRF ray-marching hot path on B200.

Contract (see the task brief): ``python bench.py --gpus N --steps K --warmup W`` prints ONE JSON
line.  A step = one full 800x800 synthetic Blender-'chair' view (640 000 rays, 64 coarse + 128
fine samples, two 8x256 intrinsic MLPs) rendered per GPU (weak scaling: every rank renders its own
pose; rays shard by image, there is no data-path collective).

  value      rays/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        rays/s through the public API (object_level.render with host rays): pinned-host
             rays -> H2D -> render -> D2H of the six maps, every step
  roofline   tensor-pipe fraction of the dominant kernel (k_mlp_tc, fine pass launch), algorithmic
             FLOPs (1 318 912 FLOP/sample) / CUDA-event time / measured bf16 peak
  cpu_baseline  the CPU oracle (torch fp32 port of the reference algorithm) on a bounded ray sample

``--impl reference`` times the reference algorithm's CPU implementation (the oracle port: the
reference itself lives in /root/reference, which does not exist on the GPU box) on the host cores.
"""
import argparse
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


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


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


def synthetic_rays(rank, device):
    """[H*W, 11] rays of a Blender-style view: pose_spherical(theta_rank, -30, 4) (load_blender.py:29-34),
    camera_angle_x of the 'chair' scene, near 2, far 6 - generated on the device by inrf_get_rays."""
    from intrinsicnerf_b200 import ops
    theta = math.radians(-180.0 + 360.0 * (rank % 100) / 100.0)
    phi = math.radians(-30.0)
    ct, st, cp, sp = math.cos(theta), math.sin(theta), math.cos(phi), math.sin(phi)
    trans = torch.tensor([[1., 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 4.0], [0, 0, 0, 1]])
    rphi = torch.tensor([[1., 0, 0, 0], [0, cp, -sp, 0], [0, sp, cp, 0], [0, 0, 0, 1]])
    rth = torch.tensor([[ct, 0, -st, 0], [0, 1., 0, 0], [st, 0, ct, 0], [0, 0, 0, 1]])
    flip = torch.tensor([[-1., 0, 0, 0], [0, 0, 1, 0], [0, 1, 0, 0], [0, 0, 0, 1]])
    c2w = (flip @ rth @ rphi @ trans)[:3, :4]
    f = 0.5 * W / math.tan(0.5 * 0.6911112070083618)
    K = [[f, 0.0, 0.5 * W], [0.0, f, 0.5 * H], [0.0, 0.0, 1.0]]
    return ops.get_rays_packed(H, W, K, c2w, 2.0, 6.0, device)


def cpu_baseline(n_rays=2048, seconds_cap=40.0):
    """The oracle's render_rays on the host cores.  The thread count is probed (all cores, then
    halving) on a small batch and the fastest setting is used for the timed sample - many-core hosts
    are slower with every core on these GEMM sizes."""
    from oracle import nerf_oracle as orc
    ncpu = os.cpu_count() or 1
    coarse, fine = orc.seeded_nets("object")
    rays = orc.blender_rays(64, 64)[:n_rays].contiguous()
    probe = {}
    with torch.no_grad():
        t = ncpu
        while t >= 8 or t == ncpu:
            torch.set_num_threads(t)
            orc.render_rays(rays[:128], coarse, fine, white_bkgd=True)
            t0 = time.perf_counter()
            orc.render_rays(rays[:256], coarse, fine, white_bkgd=True)
            probe[t] = 256 / (time.perf_counter() - t0)
            if t <= 8:
                break
            t //= 2
        threads = max(probe, key=probe.get)
        torch.set_num_threads(threads)
        orc.render_rays(rays[:256], coarse, fine, white_bkgd=True)          # warm-up
        t0 = time.perf_counter()
        done = 0
        for i in range(0, rays.shape[0], 512):
            orc.render_rays(rays[i:i + 512], coarse, fine, white_bkgd=True)
            done += min(512, rays.shape[0] - i)
            if time.perf_counter() - t0 > seconds_cap:
                break
        dt = time.perf_counter() - t0
    return {"value": done / dt, "unit": "rays/s", "cores": threads, "kind": "port", "host_cpus": ncpu,
            "thread_probe_rays_per_s": {str(k): round(v, 1) for k, v in probe.items()},
            "sample": f"{done} rays of a 64x64 Blender view, 64+128 samples, oracle.render_rays, torch {torch.__version__} CPU"}


def run_reference(args, rank, world):
    """--impl reference: the reference algorithm's CPU implementation (oracle port), rank 0 only."""
    if rank != 0:
        return
    steps = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(n_rays=1024, seconds_cap=30.0)
        if i >= args.warmup:
            steps.append(base["value"])
    v = sum(steps) / len(steps)
    base["value"] = v
    out = {"impl": "reference", "metric": "rays/sec (64+128 samples)", "value": v, "unit": "rays/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * 1024 / v, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": "blender_chair_800x800_64+128", "n_samples": N_SAMPLES, "n_importance": N_IMPORTANCE,
                      "net": "2 x NeRF(D=8,W=256,skips=[4]) intrinsic heads",
                      "sample": "bounded 1024-ray sample of the workload per step (cost is exactly linear in rays)"},
           "cpu_baseline": base, "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def run_room0(args, rank, local, world, dev, dist):
    """Secondary line (not the headline): the SSR fork's renderer on Replica-shaped input.  One step = 8 frames of
    320x240 (614 400 rays) through SSRRenderer-equivalent calls; `value` device-resident via inrf_render_fwd,
    `e2e` through SSRRenderer.render_rays with host rays and D2H of rgb / depth / semantic logits."""
    import intrinsicnerf_b200 as inrf
    from intrinsicnerf_b200 import ops, ssr
    C, Hh, Ww, frames = 28, 240, 320, 8
    flop_per_ray = (N_SAMPLES + N_SAMPLES + N_IMPORTANCE) * 2 * (692224 + 128 * C)          # SURVEY 8d, SSR network
    torch.manual_seed(20220414)
    mk = lambda: inrf.Semantic_NeRF(True, C, D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27,  # noqa: E731
                                    use_viewdirs=True).to(dev)
    coarse, fine = mk(), mk()
    poses = torch.eye(4).repeat(frames, 1, 1)
    for i in range(frames):
        a = math.radians(45.0 * i + 5.0 * rank)
        poses[i, 0, 0], poses[i, 0, 2], poses[i, 2, 0], poses[i, 2, 2] = math.cos(a), math.sin(a), -math.sin(a), math.cos(a)
    rays_dev = ssr.create_rays(frames, poses, Hh, Ww, Ww / 2, Ww / 2, (Ww - 1) / 2, (Hh - 1) / 2, 0.1, 10.0).reshape(-1, 11).contiguous()
    n_rays = rays_dev.shape[0]
    rays_host = rays_dev.cpu().pin_memory()
    pc, pf = coarse.packed(), fine.packed()

    class T(ssr.SSRRenderer):
        pass
    t = T()
    t.N_samples, t.N_importance, t.perturb, t.raw_noise_std, t.training = N_SAMPLES, N_IMPORTANCE, 0.0, 0.0, False
    t.white_bkgd, t.enable_semantic, t.num_valid_semantic_class, t.endpoint_feat = False, True, C, False
    t.netchunk = t.chunk = 76800
    t.ssr_net_coarse, t.ssr_net_fine = coarse, fine
    t.embed_fn, _ = ssr.get_embedder(10, 0, scalar_factor=10)
    t.embeddirs_fn, _ = ssr.get_embedder(4, 0, scalar_factor=1)
    chunks = [(i, min(i + 76800, n_rays)) for i in range(0, n_rays, 76800)]

    def step_device():
        for a_, b_ in chunks:
            ops.render_chunk(rays_dev[a_:b_], pc, pf, variant=1, n_classes=C, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE,
                             white_bkgd=False, pe_scalar_factor=10.0)

    out_host = [torch.empty(n_rays, c, pin_memory=True) for c in (3, 1, C)]

    def step_e2e():
        r = rays_host.to(dev, non_blocking=True)
        with torch.no_grad():
            d = t.render_rays(r)
        for dst, k in zip(out_host, ("rgb_fine", "depth_fine", "sem_logits_fine")):
            dst.copy_(d[k].reshape(n_rays, -1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s_, e_ in ev:
            s_.record()
            fn()
            e_.record()
        barrier()
        tt = torch.tensor([sum(s_.elapsed_time(e_) for s_, e_ in ev)], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms_dev = timed(step_device, args.steps, max(3, args.warmup))
        clocks = sampler.stop() if rank == 0 else None
        e2e_steps = max(2, args.steps // 2)
        ms_e2e = timed(step_e2e, e2e_steps, 1)
    if rank == 0:
        pk = peaks()
        value = n_rays * world * args.steps / (ms_dev * 1e-3)
        out = {"metric": "rays/sec (64+128 samples)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
               "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f16 operands (RN) x f32 accumulate on tcgen05; everything else f32",
               "data": "synthetic (random-init weights, reference seed 20220414; Replica-shaped cameras)",
               "config": {"workload": "replica_room0_320x240_64+128_C28", "frames_per_step_per_gpu": frames, "rays_per_step_per_gpu": n_rays,
                          "n_samples": N_SAMPLES, "n_importance": N_IMPORTANCE, "net": "2 x Semantic_NeRF(D=8,W=256,skips=[4], C=28)",
                          "parallelism": f"rays sharded by image, dp{world}, no data-path collective",
                          "l2": "no explicit flush: per-step working set >> 126 MB L2"},
               "e2e": {"value": n_rays * world * e2e_steps / (ms_e2e * 1e-3), "unit": "rays/s", "h2d_bytes_per_step": int(n_rays * 44),
                       "d2h_bytes_per_step": int(n_rays * (4 + C) * 4),
                       "api": "SSRRenderer.render_rays(host rays) + D2H of rgb_fine, depth_fine, sem_logits_fine (raw_* materialised as the API requires)"},
               "gpu_launches": 8 * len(chunks) * args.steps,
               "roofline": {"bound": "tensor", "kernel": "k_mlp_tc (whole step)", "achieved": value / world * flop_per_ray / 1e12,
                            "peak": pk["bf16_sustained"] or pk["bf16_tflops"], "unit": "TFLOP/s",
                            "frac": value / world * flop_per_ray / 1e12 / (pk["bf16_sustained"] or pk["bf16_tflops"]), "traffic": None,
                            "peak_source": pk["source"] + ", sustained bf16 (whole step)"},
               "clocks": clocks, "cpu_baseline": None}
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="tc", choices=["tc", "fp32"])
    ap.add_argument("--chunk", type=int, default=160000, help="rays per inrf_render_fwd call (bounds the raw scratch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--workload", default="chair", choices=["chair", "room0"],
                    help="chair: the BASELINE headline config (default). room0: BASELINE config 3 - Replica room_0 shape "
                         "(320x240 frames, Semantic_NeRF with 28 classes, near 0.1 / far 10, PE scale 10), 8 frames per step")
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
    from intrinsicnerf_b200 import object_level as ol, ops
    ops.set_default_precision(args.precision)

    if args.workload == "room0":
        return run_room0(args, rank, local, world, dev, dist)
    # random-init weights of the reference architecture, reference seed (run_nerf.py:1130)
    torch.manual_seed(20220414)
    mk = lambda: inrf.NeRF(D=8, W=256, input_ch=63, output_ch=5, skips=[4], input_ch_views=27, use_viewdirs=True).to(dev)  # noqa: E731
    coarse, fine = mk(), mk()
    embed_fn, _ = ol.get_embedder(10, 0)
    embeddirs_fn, _ = ol.get_embedder(4, 0)
    kw = dict(network_fn=coarse, network_fine=fine, network_query_fn=ol._FusedQuery(embed_fn, embeddirs_fn, 65536),
              N_samples=N_SAMPLES, N_importance=N_IMPORTANCE, perturb=0., white_bkgd=True, raw_noise_std=0.)
    rays_dev = synthetic_rays(rank, dev)
    rays_host = rays_dev.cpu().pin_memory()
    n_rays = rays_dev.shape[0]
    pc, pf = coarse.packed(), fine.packed()
    chunks = [(i, min(i + args.chunk, n_rays)) for i in range(0, n_rays, args.chunk)]
    launches_per_step = 8 * len(chunks)

    def step_device():
        for a, b in chunks:
            ops.render_chunk(rays_dev[a:b], pc, pf, white_bkgd=True, n_samples=N_SAMPLES, n_importance=N_IMPORTANCE)

    o_host, d_host = rays_host[:, 0:3].contiguous().pin_memory(), rays_host[:, 3:6].contiguous().pin_memory()
    out_host = [torch.empty(n_rays, c, pin_memory=True) for c in (3, 1, 1, 3, 1, 3)]
    Kmat = [[1111.111, 0, W / 2], [0, 1111.111, H / 2], [0, 0, 1]]

    def step_e2e():
        ro, rd = o_host.to(dev, non_blocking=True), d_host.to(dev, non_blocking=True)
        with torch.no_grad():
            res = ol.render(H, W, Kmat, chunk=args.chunk, rays=(ro, rd), ndc=False, near=2., far=6., use_viewdirs=True, **kw)
        for dst, src in zip(out_host, res[:6]):
            dst.copy_(src.reshape(n_rays, -1), non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for s, e in ev:
            s.record()
            fn()
            e.record()
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in ev)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    with torch.no_grad():
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ms_dev = timed(step_device, args.steps, max(3, args.warmup))
        clocks = sampler.stop() if rank == 0 else None
        ms_e2e = timed(step_e2e, max(2, args.steps // 2), 1)
        e2e_steps = max(2, args.steps // 2)

        # dominant kernel alone: fine-pass launch of k_mlp_tc (or the fp32 kernel) on one chunk
        a, b = chunks[0]
        o = ops.render_chunk(rays_dev[a:b], pc, pf, white_bkgd=True, want_z=True)
        zf = o["z_fine"]
        n_k = 5
        for _ in range(2):
            ops.mlp_forward_rays(pf, 0, 0, rays_dev[a:b], zf)
        torch.cuda.synchronize()
        ks, ke = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ks.record()
        for _ in range(n_k):
            ops.mlp_forward_rays(pf, 0, 0, rays_dev[a:b], zf)
        ke.record()
        torch.cuda.synchronize()
        k_ms = ks.elapsed_time(ke) / n_k
        k_flops = (b - a) * (N_SAMPLES + N_IMPORTANCE) * FLOP_PER_SAMPLE

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    pk = peaks()
    total_rays = n_rays * world * args.steps
    value = total_rays / (ms_dev * 1e-3)
    e2e_val = n_rays * world * e2e_steps / (ms_e2e * 1e-3)
    ach_tf = k_flops / (k_ms * 1e-3) / 1e12
    kernel_name = "k_mlp_tc" if args.precision == "tc" else "k_mlp_fp32"
    out = {
        "metric": "rays/sec (64+128 samples)", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f16 operands (RN) x f32 accumulate on tcgen05; sigma head, PE, compositing, resampling in f32" if args.precision == "tc" else "f32",
        "data": "synthetic (random-init weights, reference seed 20220414; pose_spherical Blender camera)",
        "config": {"workload": "blender_chair_800x800_64+128", "rays_per_step_per_gpu": n_rays, "n_samples": N_SAMPLES,
                   "n_importance": N_IMPORTANCE, "net": "2 x NeRF(D=8,W=256,skips=[4]) intrinsic heads", "chunk_rays": args.chunk,
                   "parallelism": f"rays sharded by image, dp{world}, no data-path collective",
                   "l2": "no explicit flush: the per-step working set (raw tensors, ~5.9 GB) is >> the 126 MB L2"},
        "e2e": {"value": e2e_val, "unit": "rays/s", "h2d_bytes_per_step": int(n_rays * 24), "d2h_bytes_per_step": int(n_rays * 48),
                "api": "object_level.render(rays=(o,d) from pinned host) + D2H of rgb,disp,acc,albedo,shading,residual"},
        "gpu_launches": launches_per_step * args.steps,
        "roofline": {"bound": "tensor", "kernel": kernel_name, "achieved": ach_tf, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                     "frac": ach_tf / pk["bf16_tflops"],
                     # DRAM bytes per launch: ncu --set full of the same kernel (profiles/r01e_prof_mlp_tc_raw.md):
                     # (35.1 MB read + 282.7 MB written) / 7 680 000 rows = 41.4 B/row vs 48 B/row algorithmic (44 raw + 4 z)
                     "traffic": (b - a) * (N_SAMPLES + N_IMPORTANCE) * 41.4, "peak_source": pk["source"] + ", burst bf16 (kernel timed alone)",
                     "kernel_ms": k_ms, "launch_rows": (b - a) * (N_SAMPLES + N_IMPORTANCE),
                     "whole_step_frac": value / world * FLOP_PER_RAY / 1e12 / (pk["bf16_sustained"] or pk["bf16_tflops"]),
                     "hbm_frac_algorithmic": value / world * BYTES_PER_RAY / 1e9 / pk["hbm_gbs"]},
        "clocks": clocks,
    }
    if not args.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_baseline()
    elif not args.no_cpu_baseline:
        out["cpu_baseline"] = None
    print(json.dumps(out), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
