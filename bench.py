#!/usr/bin/env python
"""bench.py -- rays/s (64 samples/ray) of the BTS render hot path on N B200s (BASELINE.json metric).

    python bench.py --gpus 1 --steps 20 --warmup 5              # our CUDA path (one JSON line on stdout)
    python bench.py --impl reference --steps 3 --warmup 1       # the reference algorithm on the host CPU (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...       # one rank per GPU, ray-sharded, weak scaling

Workload at N=1 = BASELINE config 2 ("KITTI-360 eval_depth: 192x640, 64 samples/ray"): one stereo frame as
``models/bts/evaluator.py:62-68`` renders it -- rays of both views (245 760 rays), encoder view 0, one colour view,
``want_weights=True, want_alphas=True``.  A step = one pass of the hot path over that ray batch: fold of the frame's
feature map + its operand-block copy (our per-frame glue), colour packing and the fused render launch (band kernel +
the work-list pass of the gather kernel).  ``encode`` (the ResNet) is excluded from the metric as in BASELINE.md.  At N>1
every rank renders its own block of an N-times larger global ray batch (``behindthescenes_b200.rays.shard_rays``; scene
replicated; weak scaling) and the scalar metric (depth sum, ray count) is accumulated on the device on a side stream and
all-reduced over NCCL once, off the critical path (the reference's ``sync_all_reduce`` runs once per epoch too,
``utils/metrics.py:31``).

Extra keys of the same JSON line (``"extra"``, skipped with ``--no-extras``): strong scaling of ONE cfg2 frame and of the
cfg5 ray batch over the ranks, the cfg3 training step (weak), the cfg4 render -- ``tools/bench_workloads.py``.

Timing: per-step CUDA events on the launching stream (torch's current stream), an L2 flush (256 MiB write) before
every timed step outside the event pair, max over ranks.  See DESIGN.md "Measurement".
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from behindthescenes_b200 import synthetic as syn  # noqa: E402

METRIC = "rays/s (64 samples/ray)"
K_SAMPLES = 64
# SURVEY.md §8d algorithmic figures for config 2 outputs (rgb, depth, invalid, weights, alphas; one ray per pixel)
BYTES_PER_RAY = 32 + 4 * 64 + 12 + 4 + 12 + 3 * 4 * K_SAMPLES      # = 1084 B
DENSE_FLOP_PER_RAY = K_SAMPLES * (2 * (64 + 39) * 64 + 2 * 64)       # = 0.852 MFLOP (MLP 103->64->1)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured (MEASURED_PEAKS.json; sustained bf16)"}
    return {"hbm_gbs": 6650.0, "tflops": 1590.0, "source": "fallback (B200_PROFILING.md)"}


def make_workload(args):
    wl = syn.cfg2_kitti360_eval(H=192, W=640, K=K_SAMPLES, stereo=True, learn_empty=True, lin_out_bias=-4.0, seed=0)
    return wl


def weights_state_dict(wl):
    net, _ = syn.build_model(wl, "cpu")
    return {k: v.clone() for k, v in net.state_dict().items() if not k.startswith("encoder.")}


# ---------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference algorithm (oracle port) on the host cores
# ---------------------------------------------------------------------------------------------------------------------

def cpu_reference_run(wl, sd, n_rays, steps, warmup):
    from oracle import bts_oracle as orc
    from tests.golden_io import mlp_params
    st = orc.encode_state(wl.images, wl.Ks, wl.poses_c2w, wl.feats, wl.ids_encoder, wl.ids_render)
    fcfg = orc.FieldConfig.from_model_conf(wl.model_conf)
    rcfg = orc.RenderConfig.from_renderer_conf(wl.renderer_conf)
    rays = wl.rays[:, :n_rays].contiguous()
    u = wl.u[:n_rays].contiguous()
    mlp = mlp_params(sd)
    out = None
    times = []
    # "all the host threads it can use": torch's intra-op pool at the core count this process may run on, unless a
    # smaller pool is faster for this op chain (many small elementwise ops) -- probe a few sizes on a 2048-ray sample.
    try:
        avail = len(os.sched_getaffinity(0))
    except Exception:
        avail = os.cpu_count() or 1
    best = (None, float("inf"))
    with torch.no_grad():
        for t in sorted({min(avail, c) for c in (8, 16, 32, 64, avail)}):
            torch.set_num_threads(t)
            r2, u2 = rays[:, :2048].contiguous(), u[:2048].contiguous()
            orc.render(st, mlp, fcfg, rcfg, r2, u=u2, empty_feature=sd.get("empty_feature"))
            t0 = time.perf_counter()
            orc.render(st, mlp, fcfg, rcfg, r2, u=u2, empty_feature=sd.get("empty_feature"))
            dt = time.perf_counter() - t0
            if dt < best[1]:
                best = (t, dt)
    torch.set_num_threads(best[0])
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            out = orc.render(st, mlp, fcfg, rcfg, rays, u=u, empty_feature=sd.get("empty_feature"),
                             want_weights=True, want_alphas=True)["coarse"]
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return sum(times) / len(times), out


def gpu_eager_reference_run(wl, sd, dev, steps=3, warmup=1):
    """The reference's own render path as it would run on this GPU: the oracle's torch op chain (same ops, same
    100000-point chunking as nerf.py:244-268) on CUDA tensors, fp32 (TF32 matmul off = torch default).  Baseline only."""
    from oracle import bts_oracle as orc
    from tests.golden_io import mlp_params
    mv = lambda t: t.to(dev)  # noqa: E731
    st = orc.encode_state(mv(wl.images), mv(wl.Ks), mv(wl.poses_c2w), mv(wl.feats), wl.ids_encoder, wl.ids_render)
    fcfg = orc.FieldConfig.from_model_conf(wl.model_conf)
    rcfg = orc.RenderConfig.from_renderer_conf(wl.renderer_conf)
    mlp = {k: mv(v) for k, v in mlp_params(sd).items()}
    ef = mv(sd["empty_feature"]) if "empty_feature" in sd else None
    rays, u = mv(wl.rays), mv(wl.u)
    evs = []
    with torch.no_grad():
        for i in range(warmup + steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = orc.render(st, mlp, fcfg, rcfg, rays, u=u, empty_feature=ef, want_weights=True, want_alphas=True)["coarse"]
            b.record()
            if i >= warmup:
                evs.append((a, b))
        torch.cuda.synchronize(dev)
    ms = sum(a.elapsed_time(b) for a, b in evs) / len(evs)
    return ms, out


def run_reference_arm(args):
    """The reference algorithm (oracle port) on the host cores, on OUR arm's config: the whole 245 760-ray stereo frame per
    step.  A frame takes ~10 s of CPU, so the number of timed steps is capped to what fits ~2 minutes (reported)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = make_workload(args)
    sd = weights_state_dict(wl)
    n_rays = wl.rays.shape[1]
    t0 = time.perf_counter()
    sec_probe, _ = cpu_reference_run(wl, sd, 15360, 1, 0)               # 1/16 of the frame: how long will a frame take?
    est = sec_probe * n_rays / 15360
    steps = max(1, min(args.steps, int(120.0 / max(est, 1e-3))))
    warm = 1 if (args.warmup > 0 and est < 30.0) else 0
    sec, _ = cpu_reference_run(wl, sd, n_rays, steps, warm)
    v = n_rays / sec
    cores = torch.get_num_threads()
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "rays/s", "n_gpus": args.gpus, "steps": steps,
            "steps_requested": args.steps, "warmup": warm, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded frustum, randn features in place of the encoder output)",
            "config": {"workload": "cfg2 KITTI-360 eval_depth: 192x640 stereo frame, 245760 rays/GPU, K=64, outputs rgb+depth+invalid+weights+alphas",
                       "rays_per_gpu": n_rays, "samples_per_ray": K_SAMPLES, "impl_note": "reference algorithm (oracle port) on the host CPU; "
                       "timed steps capped to ~2 minutes of CPU"},
            "cpu_baseline": {"value": v, "unit": "rays/s", "cores": cores, "kind": "port",
                             "sample": f"the whole frame ({n_rays} rays, K=64) per step, {steps} timed step(s); torch CPU ops chunked by 100000 points like nerf.py:244-268"},
            "e2e": {"value": v, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------------------

class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def n_samples(self):
        try:
            self.f.flush()
            return sum(1 for _ in open(self.f.name))
        except Exception:
            return 0

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(",") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for nme, val in zip(names, r[4:8]):
                    if val.strip().lower() == "active":
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        # "under load" = samples in the upper half of the power range
        thr = (max(pw) + min(pw)) / 2
        load = sorted(s for s, w in zip(sm, pw) if w >= thr) or sorted(sm)
        return {"sm_mhz": load[len(load) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw)}


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the render path has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from behindthescenes_b200 import _lib
    from behindthescenes_b200 import rays as R
    lib = _lib.load()

    wl = make_workload(args)
    sd = weights_state_dict(wl)
    net, wrapped = syn.build_model(wl, dev, state_dict=sd)
    wrapped.renderer.precision = args.precision
    frame_rays = wl.rays                                              # (1, 245760, 8) host
    n_frame = frame_rays.shape[1]
    # global ray batch = `world` frames' worth, sharded contiguously; every rank holds the replicated scene
    lo, hi = R.shard_bounds(n_frame * world, rank, world)
    rays_host = frame_rays.repeat(1, world, 1)[:, lo:hi].contiguous().pin_memory()
    rays_dev = rays_host.to(dev)
    n_rays = rays_dev.shape[1]
    with torch.no_grad():
        net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    wants = dict(want_weights=True, want_alphas=True)
    # per-step scalar metric (depth sum; ray count): accumulated on the device on a side stream, all-reduced ONCE at the
    # end -- nothing of it sits on the critical path of a step (r1: a blocking sum + clone + 8-byte NCCL all-reduce per
    # 1.6 ms step cost 10 % at 8 GPUs)
    metric_acc = torch.zeros(2, device=dev, dtype=torch.float64)
    side = torch.cuda.Stream(dev)

    def frame_step():
        net.invalidate_scene()                                         # a new frame: fold + pack run again
        net._scene_struct()
        return wrapped(rays_dev, **wants)["coarse"]

    graph = {"replay": None, "out": None, "read": None}

    def one_step(timed, ev=None):
        flush.fill_(1)                                                 # evict L2 (126 MB) -- outside the event pair
        if graph["replay"] is not None:                                # the whole step (6 launches) as ONE graph launch
            if graph["read"] is not None:                              # the replay overwrites the outputs the side stream
                torch.cuda.current_stream(dev).wait_event(graph["read"])   # summed (done long ago: an ordering, not a wait)
            if ev:
                ev[0].record()
            graph["replay"]()
            out = graph["out"]
        else:
            net.invalidate_scene()
            if ev:
                ev[0].record()
            net._scene_struct()
            if ev:
                ev[1].record()
            out = wrapped(rays_dev, **wants)["coarse"]
        done = ev[2] if ev else torch.cuda.Event()
        done.record()                                                  # end of the step on the compute stream
        side.wait_event(done)
        with torch.cuda.stream(side):
            metric_acc[0] += out["depth"].sum(dtype=torch.float64)
            metric_acc[1] += float(n_rays)
            if graph["replay"] is not None:
                graph["read"] = torch.cuda.Event()
                graph["read"].record(side)
        out["depth"].record_stream(side)
        return out

    def finish_metric():
        """the path's only collective: SUM of the accumulated scalars (NCCL, 16 bytes), once, on the side stream"""
        with torch.cuda.stream(side):
            red = metric_acc.clone()
            if world > 1:
                dist.all_reduce(red, op=dist.ReduceOp.SUM)
            fin = torch.cuda.Event(enable_timing=True)
            fin.record(side)
        return red, fin

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # nvidia-smi is started first: its start-up (NVML init) can stall CUDA calls of this process for milliseconds, so it
    # must be past that before anything is timed; afterwards it only polls.  GPU clocks ramp from idle (120 MHz) over the
    # first ~100 ms of load: bring the device to its steady state before timing.
    sampler = ClockSampler(local) if rank == 0 else None
    # A full Python garbage collection (hundreds of ms with torch loaded) in the middle of the timed loop starves the
    # GPU queue and shows up as one 50-200 ms "step": collect now, then keep the collector off while timing.
    gc.collect()
    gc.freeze()
    gc.disable()
    with torch.no_grad():
        t_ramp = time.perf_counter()
        while True:
            wrapped(rays_dev, **wants)
            torch.cuda.synchronize(dev)
            dt = time.perf_counter() - t_ramp
            # nvidia-smi needs 1-2 s to come up on a fresh box and its first NVML query stalls the running context for
            # tens of ms: do not start timing before it has delivered samples (give up waiting after 15 s)
            ready = sampler is None or sampler.p is None or sampler.n_samples() >= 3
            if (dt >= 1.5 and ready) or dt >= 15.0:
                break
    # ---- dominant kernel alone (roofline): the fused render launch, scene cached, events around the launch only ----
    with torch.no_grad():
        kev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
        for i in range(args.steps):
            flush.fill_(1)
            kev[i][0].record()
            wrapped(rays_dev, **wants)
            kev[i][1].record()
        torch.cuda.synchronize(dev)
    kern_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps

    with torch.no_grad():
        # eager steps first: they give the fold+pack / render split of a step (events between the launches) and bring the
        # caching allocator to its steady state
        evw = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(max(args.warmup, 3))]
        for e in evw:
            out = one_step(False, e)
        finish_metric()
        sync_all()
        med = lambda v: sorted(v)[len(v) // 2]        # the first eager steps after the kernel loop still grow the allocator
        prep_ms = med([e[0].elapsed_time(e[1]) for e in evw])
        render_ms = med([e[1].elapsed_time(e[2]) for e in evw])
        eager_step_ms = med([e[0].elapsed_time(e[2]) for e in evw])
        if not args.no_graph:
            # the timed step is one CUDA-graph launch (fold, fold_empty, pack_g16, pack_images, render_band, work-list
            # pass): no host-side gaps between the six launches, whatever the box's CPU load (8 ranks share the host)
            try:
                st = torch.cuda.Stream(dev)
                st.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(st):
                    frame_step()
                torch.cuda.current_stream(dev).wait_stream(st)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    graph["out"] = frame_step()
                graph["replay"] = g.replay
            except Exception as exc:   # capture is an optimisation of the loop, never a requirement
                print(f"bench: CUDA-graph capture failed ({exc}); timing eager steps", file=sys.stderr)
                torch.cuda.synchronize(dev)
                graph["replay"] = None
        for _ in range(max(args.warmup, 3)):
            out = one_step(False)              # warm-up of the timed form of the step
        finish_metric()
        sync_all()
        metric_acc.zero_()
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
        t_wall0 = time.perf_counter()
        for i in range(args.steps):
            out = one_step(True, evs[i])
        metric, fin = finish_metric()
        sync_all()
        wall = time.perf_counter() - t_wall0
    # what the metric adds to the job: the part of the side-stream work that outlasts the last step
    metric_tail_ms = max(0.0, evs[-1][2].elapsed_time(fin))
    per_step = sorted(e[0].elapsed_time(e[2]) for e in evs)
    mem = torch.cuda.memory_stats(dev)
    diag = {"step_ms_min": per_step[0], "step_ms_median": per_step[len(per_step) // 2], "step_ms_max": per_step[-1],
            "first_steps_ms": [round(e[0].elapsed_time(e[2]), 3) for e in evs[:6]], "ms_per_step_eager": eager_step_ms,
            "cuda_graph": graph["replay"] is not None,
            "cuda_mallocs": mem.get("num_device_alloc", -1), "alloc_retries": mem.get("num_alloc_retries", -1)}
    step_ms = (sum(e[0].elapsed_time(e[2]) for e in evs) + metric_tail_ms) / args.steps
    handed = int(net._band_work[0].item()) if getattr(net, "_band_work", None) is not None else None
    kernel_name = lib.bts_last_kernel().decode()

    # ---- end to end through the public API with HOST buffers -----------------------------------------------------------
    # What the reference's evaluate_step moves to the device per frame (utils/base_evaluator.py:132): images, poses,
    # intrinsics.  The step: H2D of those (pinned), colour packing + fold of the frame's features (device resident, as
    # produced by the encoder), rays generated in the kernel from the poses (render_views), D2H of the per-ray result
    # (z-depth + rgb) -- per-sample tensors stay on the device like in the reference (consumed by its loss/metrics).
    H_img, W_img = wl.images.shape[-2:]
    n_views = n_rays // (H_img * W_img)
    v_frame = wl.poses_c2w.shape[1]
    img_host = wl.images.contiguous().pin_memory()
    assert n_views == v_frame, "every rank renders one stereo frame (weak scaling)"
    poses_host = wl.poses_c2w.contiguous().pin_memory()
    Ks_host = wl.Ks.contiguous().pin_memory()
    img_dev = torch.empty_like(img_host, device=dev)
    poses_dev = torch.empty_like(poses_host, device=dev)
    Ks_dev = torch.empty_like(Ks_host, device=dev)
    depth_host = torch.empty((1, n_rays), dtype=torch.float32).pin_memory()
    rgb_host = torch.empty((1, n_rays, 3), dtype=torch.float32).pin_memory()
    e2e_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    with torch.no_grad():
        for i in range(-3, args.steps):
            flush.fill_(1)
            if i >= 0:
                e2e_ev[i][0].record()
            img_dev.copy_(img_host, non_blocking=True)                   # H2D: this frame's images, poses, intrinsics
            poses_dev.copy_(poses_host, non_blocking=True)
            Ks_dev.copy_(Ks_host, non_blocking=True)
            net.grid_c_imgs = net._take_views(img_dev * 0.5 + 0.5, wl.ids_render)   # encode()'s colour cache (models_bts.py:81,133); no list index: its H2D upload syncs the stream
            net.invalidate_scene()
            o = wrapped.render_views(poses_dev, Ks_dev, H_img, W_img, wl.model_conf["z_near"], wl.model_conf["z_far"],
                                     want_depth_z=True, **wants)["coarse"]
            depth_host.copy_(o["depth_z"], non_blocking=True)            # D2H of the per-ray result
            rgb_host.copy_(o["rgb"], non_blocking=True)
            if i >= 0:
                e2e_ev[i][1].record()
            torch.cuda.synchronize(dev)
        sync_all()
    e2e_seq_ms = sum(a.elapsed_time(b) for a, b in e2e_ev) / args.steps

    # The same steps as a 3-stage stream pipeline (how a serving loop runs them): H2D of frame i+1 and D2H of frame i-1
    # overlap the render of frame i on separate streams; every step still moves its own inputs and its own result.
    # No L2 flush here: the per-step inputs rotate over 5 feature maps (5 x 31.5 MB > L2) and arrive by DMA.
    slots = 5
    feats_rot = [torch.randn_like(net.grid_f_features[0]) for _ in range(slots)]
    s_in, s_c, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    img_d = [torch.empty_like(img_dev) for _ in range(slots)]
    pos_d = [torch.empty_like(poses_dev) for _ in range(slots)]
    Ks_d = [torch.empty_like(Ks_dev) for _ in range(slots)]
    dep_h = [torch.empty_like(depth_host).pin_memory() for _ in range(slots)]
    rgb_h = [torch.empty_like(rgb_host).pin_memory() for _ in range(slots)]

    def compute_stage(j):
        """the device part of a step through the public API: colour cache of encode(), fold + pack, render_views"""
        net.grid_f_features = [feats_rot[j]]
        net.grid_c_imgs = net._take_views(img_d[j] * 0.5 + 0.5, wl.ids_render)
        net.invalidate_scene()
        return wrapped.render_views(pos_d[j], Ks_d[j], H_img, W_img, wl.model_conf["z_near"], wl.model_conf["z_far"],
                                    want_depth_z=True, **wants)["coarse"]

    # One CUDA graph per slot (its buffers are fixed): the compute stage is ONE launch per step, so the pipeline's period
    # is the GPU's and not the host's enqueue rate (~14 launches + Python per 1.1 ms step; 8 ranks share the host)
    stage_graphs = None
    if not args.no_graph:
        try:
            with torch.no_grad():
                for b in img_d + pos_d + Ks_d:
                    b.zero_()
                torch.cuda.synchronize(dev)
                gs = []
                for j in range(slots):
                    img_d[j].copy_(img_host); pos_d[j].copy_(poses_host); Ks_d[j].copy_(Ks_host)
                    s_c.wait_stream(torch.cuda.current_stream(dev))
                    with torch.cuda.stream(s_c):
                        compute_stage(j)
                    torch.cuda.synchronize(dev)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, stream=s_c):
                        o_j = compute_stage(j)
                    gs.append((g, o_j))
                stage_graphs = gs
        except Exception as exc:
            print(f"bench: e2e CUDA-graph capture failed ({exc}); eager compute stage", file=sys.stderr)
            torch.cuda.synchronize(dev)
            stage_graphs = None
    keep, ev_in, ev_c, ev_out = {}, {}, {}, {}
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_pipe = args.steps + 6
    torch.cuda.synchronize(dev)
    with torch.no_grad():
        for i in range(n_pipe):
            j = i % slots
            with torch.cuda.stream(s_in):
                if i - slots in ev_out:
                    s_in.wait_event(ev_c[i - slots])                 # slot j's device inputs were consumed
                img_d[j].copy_(img_host, non_blocking=True)
                pos_d[j].copy_(poses_host, non_blocking=True)
                Ks_d[j].copy_(Ks_host, non_blocking=True)
                ev_in[i] = torch.cuda.Event(); ev_in[i].record(s_in)
            with torch.cuda.stream(s_c):
                s_c.wait_event(ev_in[i])
                # steady-state period of the pipeline, taken on the compute stream: start of step 3 (steps 0-2 fill the
                # pipeline) to start of step 3 + steps.  Every step in between waited for its own H2D (ev_in) and the host
                # cannot run more than `slots` steps ahead of the D2H stream (ev_out.synchronize below), so transfer
                # stalls show up in this period; recording on the copy streams instead would add up to `slots` steps of
                # skew once the host no longer syncs every step.
                if i == 3:
                    t0.record(s_c)
                if i == 3 + args.steps:
                    t1.record(s_c)
                if stage_graphs is not None:
                    if i - slots in ev_out:
                        s_c.wait_event(ev_out[i - slots])            # slot j's outputs have left for the host
                    stage_graphs[j][0].replay()
                    o = stage_graphs[j][1]
                else:
                    o = compute_stage(j)
                ev_c[i] = torch.cuda.Event(); ev_c[i].record(s_c)
            keep[i] = o                                              # outputs stay alive until their D2H has been issued+done
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c[i])
                if i - slots in ev_out:
                    ev_out[i - slots].synchronize()                  # host buffer j is free again (its consumer would read it here)
                dep_h[j].copy_(o["depth_z"], non_blocking=True)
                rgb_h[j].copy_(o["rgb"], non_blocking=True)
                ev_out[i] = torch.cuda.Event(); ev_out[i].record(s_out)
            keep.pop(i - slots, None)
        torch.cuda.synchronize(dev)
        sync_all()
    e2e_ms = t0.elapsed_time(t1) / args.steps
    e2e_graph = stage_graphs is not None
    clocks = sampler.stop() if sampler else None                     # covers warm-up, timed loop, kernel loop, e2e loops
    gc.enable()
    h2d_bytes = int((img_host.numel() + poses_host.numel() + Ks_host.numel()) * 4)

    # ---- max over ranks ----------------------------------------------------------------------------------------------
    t = torch.tensor([step_ms, e2e_ms, kern_ms, prep_ms, e2e_seq_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    step_ms, e2e_ms, kern_ms_max, prep_ms, e2e_seq_ms = (float(x) for x in t.tolist())
    total_rays = n_frame * world
    value = total_rays / (step_ms * 1e-3)
    e2e_value = total_rays / (e2e_ms * 1e-3)

    line = None
    if rank == 0:
        pk = peaks()
        rps_kernel = n_rays / (kern_ms * 1e-3)
        hbm_ach = rps_kernel * BYTES_PER_RAY / 1e9
        tf_ach = rps_kernel * DENSE_FLOP_PER_RAY / 1e12
        f_h, f_t = hbm_ach / pk["hbm_gbs"], tf_ach / pk["tflops"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tp):
            traffic = (json.load(open(tp)).get(kernel_name) or {}).get("dram_bytes_per_launch")
        if f_t >= f_h:
            roof = {"bound": "tensor", "achieved": tf_ach, "peak": pk["tflops"], "unit": "TFLOP/s", "frac": f_t}
        else:
            roof = {"bound": "hbm", "achieved": hbm_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": f_h}
        roof.update({"traffic": traffic, "kernel": f"render_{kernel_name}", "kernel_ms": kern_ms, "peak_source": pk["source"],
                     "hbm_frac": f_h, "tensor_frac": f_t, "algorithmic_bytes_per_ray": BYTES_PER_RAY,
                     "algorithmic_dense_flop_per_ray": DENSE_FLOP_PER_RAY})
        # the unit that actually binds (neither HBM nor the tensor pipe): issue-slot and LSU-pipe utilisation of the
        # kernel from the committed ncu capture of this command (profiles/kernel_metrics.json), not measured live
        mp = os.path.join(ROOT, "profiles", "kernel_metrics.json")
        if os.path.isfile(mp):
            km = json.load(open(mp)).get(kernel_name) or {}
            for k_src, k_dst in (("issue_slots_busy_pct", "issue_frac"), ("lsu_data_pipe_pct", "lsu_frac"), ("tensor_pipe_pct", "tensor_pipe_frac"),
                                 ("warp_instr_per_sample", "warp_instr_per_sample"), ("source", "binding_unit_source")):
                if k_src in km:
                    roof[k_dst] = km[k_src] / 100.0 if k_src.endswith("_pct") else km[k_src]
        line = {"metric": METRIC, "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic (seeded KITTI-shaped frustum; randn features stand in for the encoder output, random-init MLP)",
                "config": {"workload": "cfg2 KITTI-360 eval_depth: 192x640 stereo frame, 245760 rays/GPU, K=64, outputs rgb+depth+invalid+weights+alphas",
                           "rays_per_gpu": n_rays, "samples_per_ray": K_SAMPLES, "parallelism": f"ray-shard x{world}",
                           "l2": "256 MiB flush before every timed step", "precision_tier": "exact-fp32" if args.precision == 0 else "fast",
                           "kernel": kernel_name, "ms_fold_pack_eager": prep_ms, "ms_render_eager": render_ms,
                           "split_note": "fold+pack / render split of the EAGER warm-up steps (events between the launches, host gaps included); the timed step is ms_per_step, the kernel alone roofline.kernel_ms",
                           "step": "one CUDA-graph launch (6 kernels)" if diag["cuda_graph"] else "eager launches",
                           "ms_metric_tail": metric_tail_ms, "metric": "depth sum + ray count accumulated on a side stream, one NCCL all-reduce at the end",
                           "band_groups_handed_to_gather_kernel": handed, "diag": diag},
                "e2e": {"value": e2e_value, "unit": "rays/s", "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": int(depth_host.numel() * 4 + rgb_host.numel() * 4), "ms_per_step": e2e_ms,
                        "how": "3-stream pipeline (H2D | fold+pack+render | D2H), every step copies its own inputs and result; compute stage = one CUDA-graph launch per step" + ("" if e2e_graph else " (capture failed: eager)") + "; steady-state period between compute starts of step 3 and step 3+steps",
                        "ms_per_step_unpipelined": e2e_seq_ms},
                # ours, per step: fold, fold_empty, pack_g16, pack_images, render_band, render_tc (work-list pass)
                "gpu_launches": args.steps * 6,
                "roofline": roof, "clocks": clocks, "wall_s_timed_loop": wall,
                "depth_sum_check": float(metric[0].item()), "rays_check": float(metric[1].item())}
        if world == 1 and not args.no_cpu_baseline:
            # reference algorithm on this box's host cores, bounded sample, + depth abs_rel of our path vs it (same u)
            n_s = 61440
            sec, ref_out = cpu_reference_run(wl, sd, n_s, 1, 1)
            with torch.no_grad():
                # the pipelined e2e loop above rotated synthetic feature maps through the net: encode the workload's frame again
                net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
                ours = wrapped(rays_dev[:, :n_s].contiguous(), u=wl.u[:n_s].to(dev), **wants)["coarse"]
            d, dr = ours["depth"].cpu(), ref_out["depth"]
            same = (ours["invalid"].cpu() == ref_out["invalid"]).flatten(2).all(-1)
            rel = (d - dr).abs() / dr
            line["cpu_baseline"] = {"value": n_s / sec, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                    "sample": f"first {n_s} rays of the frame (K=64), one timed pass after one warm-up; torch CPU ops chunked like nerf.py:244-268"}
            line["depth_abs_rel_vs_ref"] = {"all_rays": rel.mean().item(), "rays_without_knife_edge_flag_flip": rel[same].mean().item(),
                                            "max_abs_weights": (ours["weights"].cpu() - ref_out["weights"])[same].abs().max().item(),
                                            "rays": n_s, "flag_flip_rays": int((~same).sum().item())}
        if world == 1 and not args.no_ref_gpu:
            try:
                ms, ref_gpu = gpu_eager_reference_run(wl, sd, dev)
                line["ref_eager_gpu"] = {"value": n_frame / (ms * 1e-3), "unit": "rays/s", "ms_per_frame": ms,
                                         "what": "reference op chain (oracle port) in eager PyTorch on this GPU, fp32, 100000-point chunks, same frame and outputs",
                                         "speedup_ours_vs_it": value / (n_frame / (ms * 1e-3))}
            except Exception as e:  # baseline only: never fail the bench because of it
                line["ref_eager_gpu"] = {"unavailable": repr(e)[:200]}
    # ---- the other BASELINE configs + strong scaling, as extra keys of the same line -----------------------------------
    extra = {}
    if not args.no_extras:
        from tools import bench_workloads as bw
        pk = peaks()

        state = {"dead": False}

        def guarded(name, fn):
            if state["dead"]:
                extra[name] = {"skipped": "an earlier extra left the CUDA context unusable"}
                return
            try:
                t0 = time.perf_counter()
                extra[name] = fn()
                extra[name]["wall_s"] = round(time.perf_counter() - t0, 1)
            except Exception as e:           # an extra must never take the headline number down with it
                extra[name] = {"failed": repr(e)[:300]}
            try:
                torch.cuda.synchronize(dev)
                torch.cuda.empty_cache()
            except Exception as e:
                state["dead"] = True
                extra[name] = {"failed": repr(e)[:300]}
        with torch.no_grad():
            net.encode(wl.images.to(dev), wl.Ks.to(dev), wl.poses_c2w.to(dev), ids_encoder=wl.ids_encoder, ids_render=wl.ids_render)
        wrapped.renderer.precision = args.precision
        frame_dev = frame_rays.to(dev)
        guarded("frame_with_encoder_standin", lambda: bw.frame_with_encoder(wl, sd, dev, steps=min(args.steps, 10)))
        guarded("strong_cfg2_frame", lambda: bw.strong_cfg2(net, wrapped, frame_dev, dev, min(args.steps, 50), flush))
        del frame_dev
        guarded("cfg3_train_step", lambda: bw.cfg3_train_step(dev, pk, steps=min(args.steps, 10), with_reference=(world == 1 and not args.no_ref_gpu)))
        guarded("cfg4_re10k", lambda: bw.cfg4_re10k(dev, pk, steps=min(args.steps, 20)))
        guarded("strong_cfg5_stress", lambda: bw.cfg5_stress(dev, pk, 4096 * 4096, steps=2))
    if rank == 0:
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=int, default=0, help="0 exact-fp32 tier, 1 fast tier")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the eager-PyTorch-on-GPU reference leg")
    ap.add_argument("--no-graph", action="store_true", help="time eager steps instead of one CUDA-graph launch per step")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary workloads (strong scaling, cfg3/4/5)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
