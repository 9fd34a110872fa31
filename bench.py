#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config, one JSON line on stdout (rank 0).

Metric: path samples/s (one sample = one pixel x one spp iteration of the whole ReSTIR + path-tracing pipeline) and
fwd+bwd ms/step of the stage-1 training step, config C2 (`configs[1]`): synthetic 500k-triangle mesh, 800x800 rays,
spp 4, 3 bounces (direct + 2 indirect), ReSTIR temporal + spatial reuse, LBVH rebuilt every step, backward into the
envmap, shading normals (-> vertices) and kd/ks (-> vertex texture).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config C2]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (one view per rank, weak scaling,
                                                                                   one NCCL allreduce of the gradients)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "path_samples_per_sec"
UNIT = "samples/s"


def _peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampler (recipe line of /opt/skills/guides/B200_PROFILING.md)
# ----------------------------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self._t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                self.samples.append([x.strip() for x in out.strip().split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        ok = [s for s in self.samples if len(s) == 7]
        if not ok:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(s[0]) for s in ok)
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(s[3 + k].lower().startswith("active") for s in ok)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(ok[0][1]), "reasons": reasons,
                "power_w_max": max(float(s[2]) for s in ok), "samples": len(ok)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm: the CPU oracle (the reference itself has no CPU implementation and cannot be built here)
# ----------------------------------------------------------------------------------------------------------------------
def host_threads():
    """Host threads this process may use: its CPU affinity set, not OMP_NUM_THREADS (torch.distributed.run exports
    OMP_NUM_THREADS=1 into every rank, which would turn the reference arm into a single-threaded run)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def oracle_sample(cfg_name, crop, spp, threads=None):
    """One bounded sample of the workload on the host cores: LBVH build + forward spp loop on a crop x crop frame.
    Returns (samples, seconds, counters).  The OpenMP thread count is set explicitly (all host threads by default)."""
    from oracle import oracle as O, driver as D
    from mirres_restir_nerf_mesh_b200 import synth
    O.set_threads(threads or host_threads())
    cfg = synth.CONFIGS[cfg_name]
    st = _ORACLE_STATE
    if "mesh" not in st:
        st["mesh"] = synth.make_mesh(cfg)
        st["env"] = synth.envmap(*cfg["env"])
        v, f = st["mesh"]
        b = O.Bvh(v, f)
        ro, rd = synth.camera_rays(crop, crop)
        hit, t, pos, nrm, prim = O.trace(b, ro, rd)
        st["g"] = synth.gbuffer_from_hits(ro, rd, hit, t, pos, nrm)
    v, f = st["mesh"]
    counters = O.new_counters()
    t0 = time.perf_counter()
    b = O.Bvh(v, f)  # the reference rebuilds the LBVH every step (nerf/renderer.py:975)
    D.run_no_denoise(b, st["env"], st["g"], spp, crop, crop, 4242, lambda p: synth.material(p), max_bounce=cfg["max_bounce"],
                     counters=counters)
    dt = time.perf_counter() - t0
    return crop * crop * spp, dt, counters


_ORACLE_STATE = {}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mirres_restir_nerf_mesh_b200 import synth
    from oracle import oracle as O
    cfg = synth.CONFIGS[args.config]
    crop = cfg["W"]  # the full frame of the workload
    O.set_threads(host_threads())
    cores = O.threads_in_use()  # measured inside an OpenMP region of the oracle, not read from the environment
    spp_sample = _oracle_spp(args, cfg)
    for _ in range(args.warmup):
        oracle_sample(args.config, crop, spp_sample)
    tot_s, tot_t = 0, 0.0
    for _ in range(args.steps):
        s, dt, _ = oracle_sample(args.config, crop, spp_sample)
        tot_s += s
        tot_t += dt
    value = tot_s / tot_t
    sample = ("CPU oracle (oracle/, C++/OpenMP restatement of the Slang kernels; the reference has no CPU path): LBVH "
              "rebuild of the %s mesh + forward spp loop (%d of the workload's %d spp iterations, %d bounces) on a %dx%d frame, "
              "no backward" % (args.config, spp_sample, cfg["spp"], cfg["max_bounce"] + 1, crop, crop))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(args, cfg),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "omp_num_threads_env": os.environ.get("OMP_NUM_THREADS")},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def _workload_name(name, cfg, mode="train"):
    if mode == "render":
        return ("%s: novel-view render (forward only), %s mesh, %dx%d rays, spp=%d, %d bounces (direct + %d indirect), ReSTIR "
                "initial + temporal + spatial, LBVH build + G-buffer + spp loop + denoise + composite per frame" %
                (name, "x".join(str(x) for x in cfg["mesh"][1].values()), cfg["W"], cfg["H"], cfg["spp"], cfg["max_bounce"] + 1,
                 cfg["max_bounce"]))
    return ("%s: stage-1 training step, %s mesh, %dx%d rays, spp=%d, %d bounces (direct + %d indirect), ReSTIR initial + "
            "temporal + spatial, LBVH rebuild, fwd+bwd" % (name, "x".join(str(x) for x in cfg["mesh"][1].values()), cfg["W"],
                                                          cfg["H"], cfg["spp"], cfg["max_bounce"] + 1, cfg["max_bounce"]))


def _oracle_spp(args, cfg):
    """spp iterations of the bounded CPU sample: the whole loop of the training step, a few iterations of a render (the
    oracle needs ~0.2 s per iteration at 800 x 800 and ~15 s at 2048 x 2048 with 2 M triangles on 16 cores)."""
    if args.mode != "render":
        return cfg["spp"]
    return min(cfg["spp"], 8 if cfg["W"] * cfg["H"] <= 800 * 800 else 1)


def _config(args, cfg):
    """The workload both arms run -- the same dict in both JSON lines (how an arm executes it is said elsewhere:
    `execution` in the GPU line, `cpu_baseline.sample` in the reference line)."""
    if args.mode == "render":
        return {"workload": _workload_name(args.config, cfg, "render"), "l2": "256 MiB flush between timed frames",
                "parallelism": "one frame cut into row bands of equal foreground-pixel count, one band per rank; 31 halo rows "
                               "of the reservoirs received point to point per spp iteration; accumulated images all-gathered "
                               "before the denoiser"}
    return {"workload": _workload_name(args.config, cfg) + (
                "; G-buffer normals = auto_normals -> interpolation -> prepare_shading_normal, gradient to vertex "
                "positions" if getattr(args, "mesh_normals", False) else ""),
            "l2": "256 MiB flush between timed steps",
            "parallelism": "one view per rank; per step the NCCL all-reduce of the env | vertex | texture gradient buffer "
                           "(12 MB), issued inside the step in two parts"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
class ProfiledKernels:
    """Wraps Kernels: counts launches and (optionally) brackets every ABI call with CUDA events on the launching
    stream, so per-kernel device time is measured live inside the timed region."""
    # kernels per entry point (ray-casting entry points: zero fill / queue reset + gen + tracer + resolve, the bounces a
    # prologue before them); checked against the ncu launch list of a step (profiles/r9z_launches.csv.gz: 193 mr:: launches)
    LAUNCHES = {"bvh_build": 9, "env_build_distribution": 2, "eaw_bwd": 2, "workspace_prepare": 3, "initial_resampling": 4,
                "spatial_resampling": 4, "final_visibility": 4, "bounce_first": 5, "bounce_shade": 5, "gbuffer_primary": 4}

    def __init__(self, inner, torch):
        self._inner, self._torch = inner, torch
        self.launches, self.events, self.record = 0, [], False

    def __getattr__(self, name):
        fn = getattr(self._inner, name)
        if not callable(fn) or name.startswith("_") or name in ("bvh_sizes", "set_tuning", "get_tuning"):
            return fn

        def wrapped(*a, **kw):
            self.launches += self.LAUNCHES.get(name, 1)
            if self.record:
                e0 = self._torch.cuda.Event(enable_timing=True)
                e1 = self._torch.cuda.Event(enable_timing=True)
                e0.record()
                r = fn(*a, **kw)
                e1.record()
                self.events.append((name, e0, e1))
                return r
            return fn(*a, **kw)
        return wrapped

    def per_kernel_ms(self):
        out = {}
        for name, e0, e1 in self.events:
            d = out.setdefault(name, [0.0, 0])
            d[0] += e0.elapsed_time(e1)
            d[1] += 1
        return out


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R, slangpy_shim, kernels as K
    from mirres_restir_nerf_mesh_b200.graphed import CapturedStep
    from mirres_restir_nerf_mesh_b200 import meshutils as MU, renderutils_ops as OPS

    class _Interpolated(torch.autograd.Function):
        """the interpolated vertex normal mirres_gbuffer_primary has written, made differentiable w.r.t. the vertex
        normals: the backward is the barycentric scatter mirres_interpolate_bwd"""
        @staticmethod
        def forward(ctx, vnrm, smooth, prim, bary, tri, k):
            ctx.save_for_backward(prim, bary, tri)
            ctx.k, ctx.V = k, vnrm.shape[0]
            return smooth.view_as(smooth)

        @staticmethod
        def backward(ctx, g):
            prim, bary, tri = ctx.saved_tensors
            out = torch.zeros(ctx.V, 3, device=g.device)
            ctx.k.interpolate_bwd(g.contiguous(), prim, bary, tri, out)
            return out, None, None, None, None, None

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS[args.config]
    W, H, spp, mb = cfg["W"], cfg["H"], cfg["spp"], cfg["max_bounce"]
    n = W * H

    pk = ProfiledKernels(K.Kernels(), torch)
    slangpy_shim.set_kernels(pk)
    for item in args.tune or []:  # launch-shape experiments (results do not depend on them), e.g. --tune closest_split=2
        name, value = item.split("=")
        pk.set_tuning(getattr(K.Kernels, "TUNE_" + name.upper()), int(value))

    # ---- synthetic scene; each rank renders its own training view (weak scaling over views) -------------------------
    vert_np, tri_np = synth.make_mesh(cfg)
    env_np = synth.envmap(*cfg["env"])
    pose_np = synth.camera_pose(view=rank * 7)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    # the inputs of a step: mesh, envmap, camera pose (rays are generated on the device from the pose, as the reference's
    # get_rays does)
    # what changes from step to step comes from the host: vertex positions, envmap, camera pose.  The index buffer of the
    # mesh is static in stage 1 (nerf/renderer.py:967-975 moves vertices, never re-meshes): it is uploaded once
    host = dict(vert=pin(vert_np), env=pin(env_np), pose=pin(pose_np))
    device_in = {k: v.to(dev) for k, v in host.items()}
    tri_dev = torch.from_numpy(np.ascontiguousarray(tri_np)).to(dev)
    worker = R.restirbvhWorker(device_in["vert"], tri_dev)
    mat = synth.ProceduralMaterial(0.0)
    mods = R.load_m_for_restir(W, H, device=dev, max_bounce=mb)
    V, ne = vert_np.shape[0], env_np.size
    ar_stream = torch.cuda.Stream() if world > 1 else None
    collective_in_step = [world > 1 and not args.allreduce_after]
    target = torch.full((n, 3), 0.5, device=dev)
    flat_grad = torch.zeros(ne + V * 3 + V * 5, device=dev)  # env | vertex-normal | vertex-texture (kd, rough, metal)
    opts = dict(overlap=not args.no_overlap)
    bvh_stream, light_stream = torch.cuda.Stream(), torch.cuda.Stream()

    def full_step(vert, env, pose, tri=tri_dev):
        """One stage-1 training step of one view: LBVH rebuild (nerf/renderer.py:975), camera rays, G-buffer, ReSTIR +
        path tracer, denoise + composite, loss, backward into env / normals / kd / ks, scatter to vertices and vertex
        texture."""
        # three independent preambles side by side: the LBVH rebuild; the camera rays (a chain of ~20 small torch kernels,
        # 180 us -- with the lighting behind them on one stream they, not the rebuild, decided when the primary rays
        # could start, profiles/r8z_timeline_graph_replay.txt); the environment distribution + the light tiles of the
        # first spp iterations, which nothing needs before the spp loop
        cur = torch.cuda.current_stream()
        flat_grad.zero_()  # early: the fill has left the serial tail of the step by the time the scatters need it
        bvh_stream.wait_stream(cur)
        light_stream.wait_stream(cur)
        with torch.cuda.stream(bvh_stream):
            worker.update_mesh(vert, tri)
        with torch.cuda.stream(light_stream):
            env_l = env.detach().clone().requires_grad_(True)
            lighting = R.prepare_lighting(mods[0], mods[1], mods[8], mods[9], mods[10], env_l, spp, 1234 + 17 * rank,
                                          frame_pixels=n)
        rays_o, rays_d = synth.camera_rays_torch(W, H, pose)
        if args.mesh_normals:
            vert_l = vert.detach().clone().requires_grad_(True)
            vnrm, _ = MU.auto_normals(vert_l, tri)  # beside the LBVH build
        cur.wait_stream(bvh_stream)
        occ, depth = torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
        pos, nrm = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
        prim, bary = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, 2, device=dev)
        if args.mesh_normals:
            # the reference's normal chain (nerf/renderer.py:979-1030): vertex normals of the mesh -> interpolated ->
            # prepare_shading_normal, differentiable back to the vertex POSITIONS
            geom = torch.empty(n, 3, device=dev)
            pk.gbuffer_primary(worker.packed, rays_o, rays_d, occ, pos, nrm, depth, prim, bary, vnormal=vnrm.detach(),
                               tri=tri, ws=slangpy_shim.workspace(dev, n), geom_normal=geom)
            smooth = _Interpolated.apply(vnrm, nrm, prim, bary, tri, pk)
            normal = OPS.prepare_shading_normal(pos.view(1, H, W, 3), pose[3].reshape(1, 1, 1, 3), None,
                                                smooth.view(1, H, W, 3), torch.zeros(1, 1, 1, 3, device=dev),
                                                geom.view(1, H, W, 3)).view(n, 3)
        else:
            # the primary-ray trace runs alone on the GPU (everything after it needs the G-buffer): it takes a full grid
            with slangpy_shim.trace_blocks(closest_blocks=args.primary_blocks):
                pk.gbuffer_primary(worker.packed, rays_o, rays_d, occ, pos, nrm, depth, prim, bary,
                                   ws=slangpy_shim.workspace(dev, n))
            normal = nrm.requires_grad_(True)
        kd, rs = mat.gbuffer_materials(pos, occ)   # stand-in for the tiny-cuda-nn material MLP (out of scope)
        kd.requires_grad_(True)
        rs.requires_grad_(True)
        cur.wait_stream(light_stream)
        outs = R.run_restir_di_with_pt(False, 1, 1, 1, mat, None, worker, *mods, env_l, occ, normal, depth, kd, rs, rays_d,
                                       pos, None, None, None, None, W, H, spp, 2, 2, 2.0, 0.1, 0.001,
                                       random_offset=1234 + 17 * rank, max_bounce=mb, lighting=lighting, **opts)
        loss = torch.nn.functional.mse_loss(outs[0], target)
        loss.backward()
        # gradients leave the path as grad_env [He,We,3] and dense per-pixel grads; the latter are scattered to vertices /
        # vertex texture here (the reference: nvdiffrast / tcnn backward), everything lands in ONE flat buffer
        flat_grad[:ne].copy_(env_l.grad.reshape(-1))
        if collective_in_step[0]:
            # the per-step collective, part 1: the envmap segment is final here; its all-reduce runs beside the vertex scatters
            ar_stream.wait_stream(cur)
            with torch.cuda.stream(ar_stream):
                dist.all_reduce(flat_grad[:ne])
        if args.mesh_normals:
            flat_grad[ne:ne + 3 * V].copy_(vert_l.grad.reshape(-1))  # d loss / d vertex positions through the normals
        else:
            pk.interpolate_bwd(normal.grad, prim, bary, tri, flat_grad[ne:ne + 3 * V].view(V, 3))
        pk.interpolate_bwd(torch.cat((kd.grad, rs.grad), dim=1), prim, bary, tri, flat_grad[ne + 3 * V:].view(V, 5))
        if collective_in_step[0]:
            dist.all_reduce(flat_grad[ne:])  # part 2: vertex-normal and vertex-texture segments
            cur.wait_stream(ar_stream)
        return loss.detach(), flat_grad

    def finish():
        if world > 1 and not collective_in_step[0]:
            dist.all_reduce(flat_grad)  # the per-step collective after the step: texture, envmap and vertex gradients

    for _ in range(max(args.warmup, 3)):
        full_step(**device_in)
        finish()
    torch.cuda.synchronize()
    captured = None
    if not args.no_graph:
        try:
            captured = CapturedStep(full_step, device_in, warmup=1)  # NCCL all-reduces of the step are captured with it
        except Exception:
            if not collective_in_step[0]:
                raise
            # a collective that cannot be captured on this stack: one all-reduce after the replay instead (every rank
            # takes the same branch: capture fails or succeeds on all of them alike)
            torch.cuda.synchronize()
            collective_in_step[0] = False
            captured = CapturedStep(full_step, device_in, warmup=1)

    def run_step(from_host):
        if captured is not None:
            if from_host:
                captured.load(**host)
            out = captured.replay()
        else:
            src = {k: v.to(dev, non_blocking=True) for k, v in host.items()} if from_host else device_in
            out = full_step(**src)
        finish()
        return out

    if args.timeline:
        _timeline(torch, lambda _: run_step(False), args.timeline)

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)  # > 126 MB L2

    # results come back into pinned host buffers (a pageable destination makes the copy synchronous and staged)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()
    grad_host = torch.empty(ne, dtype=torch.float32).pin_memory()

    def timed(k_steps, from_host):
        total_ms, d2h = 0.0, 0
        for _ in range(k_steps):
            flush.fill_(1.0)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss, fg = run_step(from_host)
            if from_host:
                loss_host.copy_(loss, non_blocking=True)
                grad_host.copy_(fg[:ne], non_blocking=True)
                d2h = loss_host.numel() * 4 + grad_host.numel() * 4
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms, d2h

    for _ in range(args.warmup):
        run_step(False)
    torch.cuda.synchronize()
    with Clocks(local) as clocks:
        ms_dev, _ = timed(args.steps, False)
    run_step(True)
    torch.cuda.synchronize()
    ms_e2e, d2h_bytes = timed(args.steps, True)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    # per-kernel device times: an eager pass with the chains serialised (events bracket every C-ABI call on the stream
    # it is launched on; with the two chains overlapped the brackets of concurrent kernels would not be comparable)
    opts["overlap"] = False
    in_step = collective_in_step[0]
    collective_in_step[0] = False  # rank 0 runs this pass alone
    per_kernel, launches = {}, 0
    if rank == 0:
        full_step(**device_in)
        torch.cuda.synchronize()
        pk.launches, pk.events, pk.record = 0, [], True
        k_steps = 2
        for _ in range(k_steps):
            flush.fill_(1.0)
            full_step(**device_in)
        torch.cuda.synchronize()
        pk.record = False
        per_kernel, launches = pk.per_kernel_ms(), pk.launches // k_steps

    tms = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    per_rank = per_rank_alone = None
    if world > 1:
        # every rank's own time beside the maximum.  The in-step all-reduce makes the ranks wait for one another, so these
        # come out equal; what tells the imbalance between views from the cost of the collective is the second set: the
        # same step WITHOUT the collective, every rank for itself (outside the reported timing)
        allt = [torch.empty_like(tms) for _ in range(world)]
        dist.all_gather(allt, tms)
        per_rank = [[round(float(t[0]) / args.steps, 4), round(float(t[1]) / args.steps, 4)] for t in allt]
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        if in_step and captured is not None:
            opts["overlap"] = not args.no_overlap  # (the per-kernel pass above had serialised the chains)
            alone = CapturedStep(full_step, device_in, warmup=1)  # collective_in_step is off by now: no NCCL in this graph
            alone.replay()
            t_alone = 0.0
            for _ in range(max(args.steps // 2, 3)):
                flush.fill_(1.0)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                alone.replay()
                e1.record()
                torch.cuda.synchronize()
                t_alone += e0.elapsed_time(e1)
            mine = torch.tensor([t_alone / max(args.steps // 2, 3)], device=dev, dtype=torch.float64)
            alla = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(alla, mine)
            per_rank_alone = [round(float(t), 4) for t in alla]
            alone.graph.reset()
    ms_dev, ms_e2e = float(tms[0]), float(tms[1])
    samples = n * spp * world * args.steps
    value = samples / (ms_dev * 1e-3)
    e2e = samples / (ms_e2e * 1e-3)

    if rank == 0:
        peaks = _peaks()
        peak = peaks["hbm_gbs"] if peaks else 6650.0
        # foreground pixels of this rank's view, counted on the device outside the timed region: background pixels leave
        # every kernel after the compaction, so the per-pixel streams of SURVEY.md 8d are charged to the foreground only
        occ_c, depth_c = torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
        pos_c, nrm_c = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
        ro_c, rd_c = synth.camera_rays_torch(W, H, device_in["pose"])
        pk.gbuffer_primary(worker.packed, ro_c, rd_c, occ_c, pos_c, nrm_c, depth_c, ws=slangpy_shim.workspace(dev, n))
        n_fg = int((occ_c >= 0.1).sum().item())
        roof = roofline(args.config, cfg, per_kernel, 2, peak, "measured" if peaks else "fallback", n_fg)
        step_alg_all, step_alg, step_alg_ref = step_algorithmic_bytes(args.config, cfg, spp, n_fg)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": _config(args, cfg),
                "collective": None if world == 1 else ("2 all-reduces captured in the step's graph (envmap segment beside the vertex "
                                                       "scatters, then the vertex segments)" if in_step else "1 all-reduce after the step"),
                "exact_prunings": exact_prunings(),
                "execution": ("CUDA graph replay of the whole step" if captured is not None else "eager") +
                             (", concurrent schedule (reuse chain, initial candidates, shading and the indirect chains "
                              "on their own streams)" if not args.no_overlap else ""),
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes,
                        "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": launches * args.steps, "clocks": clocks.summary(), "roofline": roof,
                # S_screen charged on foreground pixels (+ 4 B of occupancy per background pixel and stage-entry); the figure
                # with S_screen charged on every pixel of the frame (round 1's) is kept beside it as *_all_pixels
                "step_roofline": None if step_alg is None else {
                    "alg_bytes_per_step": step_alg, "achieved": step_alg / (ms_dev / args.steps * 1e-3) / 1e9,
                    "peak": peak, "unit": "GB/s", "frac": step_alg / (ms_dev / args.steps * 1e-3) / 1e9 / peak,
                    "foreground_pixels": n_fg, "frame_pixels": n,
                    "alg_bytes_per_step_all_pixels": step_alg_all,
                    "frac_all_pixels": step_alg_all / (ms_dev / args.steps * 1e-3) / 1e9 / peak,
                    # traversal bytes of ALL the reference's rays, including those the product proves dead and skips
                    "frac_reference_rays": step_alg_ref / (ms_dev / args.steps * 1e-3) / 1e9 / peak},
                "kernel_ms_per_step": {k: round(v[0] / 2, 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])}}
        if per_rank is not None:
            line["per_rank_ms_per_step"] = {"device": [t[0] for t in per_rank], "e2e": [t[1] for t in per_rank]}
            if per_rank_alone is not None:
                line["per_rank_ms_per_step"]["device_without_collective"] = per_rank_alone
        if world == 1 and not args.no_cpu_baseline:
            crop = cfg["W"]
            s, dt, _ = oracle_sample(args.config, crop, spp)
            s2, dt2, _ = oracle_sample(args.config, crop, spp)
            from oracle import oracle as O
            line["cpu_baseline"] = {"value": (s + s2) / (dt + dt2), "unit": UNIT, "cores": O.threads_in_use(), "kind": "port",
                                    "sample": "CPU oracle (OpenMP, all host cores): LBVH rebuild + forward spp loop of the same %dx%d frame, 2 repetitions; no G-buffer, denoiser or backward" % (crop, crop)}
        print(json.dumps(line))
    if world > 1:
        _leave(torch, dist, captured)


def run_render(args):
    """BASELINE configs[2] / configs[4]: one frame, forward only, cut into row bands over the ranks (dist.RowBandShard).
    A step = one whole frame: LBVH build, full-frame primary G-buffer (every rank: the denoiser needs it and it is one
    trace against spp iterations), spp loop on the band, gather, denoise, composite.  Strong scaling."""
    import torch
    import torch.distributed as dist
    from mirres_restir_nerf_mesh_b200 import synth, renderer_restir as R, slangpy_shim, kernels as K, dist as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = dict(synth.CONFIGS[args.config])
    if args.spp:
        cfg["spp"] = args.spp
    W, H, spp, mb = cfg["W"], cfg["H"], cfg["spp"], cfg["max_bounce"]
    n = W * H
    pk = ProfiledKernels(K.Kernels(), torch)
    slangpy_shim.set_kernels(pk)
    for item in args.tune or []:
        name, value = item.split("=")
        pk.set_tuning(getattr(K.Kernels, "TUNE_" + name.upper()), int(value))

    vert_np, tri_np = synth.make_mesh(cfg)
    env_np = synth.envmap(*cfg["env"])
    pose_np = synth.camera_pose(view=0)  # the view tools/oracle_counters.py counts node pops and triangle tests for
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = dict(vert=pin(vert_np), tri=pin(tri_np), env=pin(env_np), pose=pin(pose_np))
    device_in = {k: v.to(dev) for k, v in host.items()}
    worker = R.restirbvhWorker(device_in["vert"], device_in["tri"])
    mat = synth.ProceduralMaterial(0.0)
    mods = R.load_m_for_restir(W, H, device=dev, max_bounce=mb)

    def gbuffer(vert, tri, pose):
        worker.update_mesh(vert, tri)
        rays_o, rays_d = synth.camera_rays_torch(W, H, pose)
        occ, depth = torch.empty(n, 1, device=dev), torch.empty(n, 1, device=dev)
        pos, nrm = torch.empty(n, 3, device=dev), torch.empty(n, 3, device=dev)
        with slangpy_shim.trace_blocks(closest_blocks=args.primary_blocks):
            pk.gbuffer_primary(worker.packed, rays_o, rays_d, occ, pos, nrm, depth, ws=slangpy_shim.workspace(dev, n))
        return rays_d, occ, depth, pos, nrm

    # the band boundaries follow the foreground of the view (one host synchronisation per view, outside the timed frames:
    # the bench renders the same view again and again)
    with torch.no_grad():
        occ0 = gbuffer(device_in["vert"], device_in["tri"], device_in["pose"])[1]
        bounds = D.balanced_bounds(occ0, W, H, world, min_rows=1) if not args.uniform_bands else D.uniform_bounds(H, world)
        n_fg = int((occ0 >= 0.1).sum().item())
    shard = D.RowBandShard(W, H, rank=rank, world=world, bounds=bounds)
    spp_full = spp

    def frame(vert, tri, env, pose):
        with torch.no_grad():
            rays_d, occ, depth, pos, nrm = gbuffer(vert, tri, pose)
            kd, rs = mat.gbuffer_materials(pos, occ)
            outs = R.run_restir_di_with_pt(False, 1, 1, 1, mat, None, worker, *mods, env, occ, nrm, depth, kd, rs, rays_d, pos,
                                           None, None, None, None, W, H, spp, 2, 2, 2.0, 0.1, 0.001, random_offset=1234,
                                           max_bounce=mb, shard=shard, overlap=not args.no_overlap)
        return outs[0]


    class _Alone(D.RowBandShard):  # a band without its neighbours: what a rank spends outside communication
        def exchange(self, tensors, row0=0):
            pass

        def gather_bands(self, images, row0=0):
            return [torch.zeros((self.fy * self.fx, im.shape[1]), device=im.device) for im in images]

    if world > 1 and not args.uniform_bands and args.rebalance > 0:
        # foreground counts do not see that some regions cast longer rays: every rank times a few iterations of ITS band
        # alone (no exchange: the values are wrong, the cost is right), the times are all-gathered and the bands cut again
        # into pieces of equal measured cost.  Part of the per-view setup (a few frames' worth of iterations), like the
        # partition itself.
        for _ in range(args.rebalance):
            shard = _Alone(W, H, rank=rank, world=world, bounds=bounds)
            took = {}
            from mirres_restir_nerf_mesh_b200.graphed import CapturedStep
            for spp in (3, 9):
                # timed as graph replays: eager launches of a band this small are bound by the host, not by the GPU
                cap = CapturedStep(frame, device_in, warmup=1)
                cap.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                cap.replay()
                cap.replay()
                e1.record()
                torch.cuda.synchronize()
                took[spp] = e0.elapsed_time(e1) / 2
                cap.graph.reset()
                del cap
            # six iterations of the band; the per-frame part (LBVH, G-buffer, denoiser) is the same on every rank and drops out
            mine = torch.tensor([took[9] - took[3]], device=dev, dtype=torch.float64)
            allt = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(allt, mine)
            bounds = D.rebalanced_bounds(bounds, [max(float(t), 1e-3) for t in allt])
        spp = spp_full
        shard = D.RowBandShard(W, H, rank=rank, world=world, bounds=bounds)
    if args.emulate:
        # diagnostics on ONE GPU: the work of rank R of W ranks, halo exchange and gather left out (images are wrong,
        # timings are what that rank would spend outside communication) -- for timelines of a band at sizes that
        # otherwise need W GPUs
        r_, w_ = (int(x) for x in args.emulate.split("/"))
        b_ = [int(x) for x in args.emulate_bounds.split(",")] if args.emulate_bounds else D.balanced_bounds(occ0, W, H, w_, min_rows=1)
        shard = _Alone(W, H, rank=r_, world=w_, bounds=b_)

    for _ in range(2):
        frame(**device_in)
    torch.cuda.synchronize()
    captured, why_eager = None, None
    if not args.no_graph:
        from mirres_restir_nerf_mesh_b200.graphed import CapturedStep
        try:
            captured = CapturedStep(frame, device_in, warmup=1)
        except Exception as e:  # e.g. a collective that cannot be captured: the frame then runs eagerly, and the line says so
            why_eager = "%s: %s" % (type(e).__name__, str(e)[:200])
            captured = None
            torch.cuda.synchronize()

    def run_step(from_host):
        if captured is not None:
            if from_host:
                captured.load(**host)
            return captured.replay()
        src = {k: v.to(dev, non_blocking=True) for k, v in host.items()} if from_host else device_in
        return frame(**src)

    if args.timeline:
        _timeline(torch, lambda _: run_step(False), args.timeline)
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    image_host = torch.empty((n, 3), dtype=torch.float32).pin_memory()

    def timed(k_steps, from_host):
        total_ms, d2h = 0.0, 0
        for _ in range(k_steps):
            flush.fill_(1.0)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            img = run_step(from_host)
            if from_host:
                if rank == 0:
                    image_host.copy_(img, non_blocking=True)  # the frame leaves through rank 0
                d2h = image_host.numel() * 4
            e1.record()
            torch.cuda.synchronize()
            total_ms += e0.elapsed_time(e1)
        return total_ms, d2h

    for _ in range(max(args.warmup, 1)):
        run_step(False)
    torch.cuda.synchronize()
    with Clocks(local) as clocks:
        ms_dev, _ = timed(args.steps, False)
    run_step(True)
    torch.cuda.synchronize()
    ms_e2e, d2h_bytes = timed(args.steps, True)
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    # per-kernel device times and the launch count: one eager frame with a few spp iterations, chains serialised
    per_kernel, launches_per_iter, launches_fixed = {}, 0, 0
    if rank == 0 or world > 1:
        spp_full = spp
        counts = []
        for spp_probe in (2, 4):
            spp = spp_probe
            pk.launches, pk.events, pk.record = 0, [], (spp_probe == 4 and rank == 0)
            saved = args.no_overlap
            args.no_overlap = True
            flush.fill_(1.0)
            frame(**device_in)
            torch.cuda.synchronize()
            args.no_overlap = saved
            counts.append(pk.launches)
        pk.record = False
        spp = spp_full
        per_kernel = pk.per_kernel_ms()
        launches_per_iter = (counts[1] - counts[0]) // 2
        launches_fixed = counts[0] - 2 * launches_per_iter

    tms = torch.tensor([ms_dev, ms_e2e], device=dev, dtype=torch.float64)
    per_rank = None
    if world > 1:
        allt = [torch.empty_like(tms) for _ in range(world)]
        dist.all_gather(allt, tms)
        per_rank = [[round(float(t[0]) / args.steps, 3), round(float(t[1]) / args.steps, 3)] for t in allt]
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e = float(tms[0]), float(tms[1])
    samples = n * spp * args.steps  # the frame is ONE job whatever the number of ranks (strong scaling)
    if rank == 0:
        peaks = _peaks()
        peak = peaks["hbm_gbs"] if peaks else 6650.0
        roof = roofline(args.config, cfg, per_kernel, 1, peak, "measured" if peaks else "fallback", n_fg, probe_spp=4)
        line = {"metric": METRIC, "value": samples / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": _config(args, cfg),
                "exact_prunings": exact_prunings(),
                "execution": ("CUDA graph replay of the whole frame (halo exchange and gather captured with it)"
                              if captured is not None else "eager launches" + (" (graph capture failed: %s)" % why_eager if why_eager else "")) +
                             (", concurrent schedule" if not args.no_overlap else ""),
                "e2e": {"value": samples / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes,
                        "d2h_bytes_per_step": d2h_bytes, "ms_per_step": ms_e2e / args.steps},
                "gpu_launches": (launches_fixed + launches_per_iter * spp) * args.steps,
                "clocks": clocks.summary(), "roofline": roof,
                "bands": {"bounds": shard.bounds, "halo_bytes_per_iteration_rank0": shard.halo_bytes(),
                          "foreground_pixels": n_fg, "frame_pixels": n},
                "kernel_ms_per_frame_spp4_probe": {k: round(v[0], 4) for k, v in sorted(per_kernel.items(), key=lambda kv: -kv[1][0])[:12]}}
        if per_rank is not None:
            line["per_rank_ms_per_step"] = {"device": [t[0] for t in per_rank], "e2e": [t[1] for t in per_rank]}
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            s_, dt, _ = oracle_sample(args.config, cfg["W"], _oracle_spp(args, cfg))
            line["cpu_baseline"] = {"value": s_ / dt, "unit": UNIT, "cores": O.threads_in_use(), "kind": "port",
                                    "sample": "CPU oracle (OpenMP, all host cores): LBVH build + %d spp iterations of the same %dx%d frame, "
                                              "forward, no G-buffer or denoiser" % (_oracle_spp(args, cfg), cfg["W"], cfg["H"])}
        print(json.dumps(line))
    if world > 1:
        _leave(torch, dist, captured)


def _leave(torch, dist, captured):
    """End of a multi-rank run.  A CUDA graph that captured NCCL work keeps the communicator busy: it is released before
    the process group goes away, and a rank that still cannot tear the group down within its barrier leaves anyway (the
    JSON line is out by then)."""
    sys.stdout.flush()
    if captured is not None:
        captured.graph.reset()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    t = threading.Timer(20.0, lambda: os._exit(0))
    t.daemon = True
    t.start()
    dist.destroy_process_group()
    t.cancel()


def exact_prunings():
    """What the product does not compute although the reference does, because the result cannot change the output (all
    bit-exact against the oracle, which computes and traces everything).  Nothing is cached across steps:
    every step rebuilds the tree, zeroes its tags and resamples from scratch."""
    from mirres_restir_nerf_mesh_b200 import renderer_restir as R
    return ["RIS target function: specular term not evaluated when the specular weight is 0 (it is multiplied by F = 0)",
            "spatial pass: visibility rays that multiply a zero target density (light not above the start surface's horizon) "
            "or only feed a weight with W = 0 are not cast",
            "final visibility: rays whose answer an earlier pass of the same spp loop produced are not cast (visibility tags: %s; "
            "MIRRES_VIS_TAGS=0 switches them off)" % ("on" if R.USE_VIS_TAGS else "off")]


def step_algorithmic_bytes(cfg_name, cfg, spp, n_foreground=None):
    """Algorithmic bytes of one whole step: sum over samples of S_screen + 36 V_n + 48 V_t, forward
    (872 B first spp iteration, 1040 B after) + backward (280 B), plus 424 B per rebuilt triangle.
    Returns (all_pixels, foreground): the first charges S_screen to every pixel of the frame, the second to the
    foreground pixels only plus the 4-byte occupancy read with which a background pixel leaves each of the table's
    11 forward + 2 backward stage entries per spp iteration (every reference kernel early-outs on occ < 0.1).  The
    traversal term is the oracle's exact node / triangle count and is the same in both."""
    path = os.path.join(ROOT, "profiles", "oracle_counters_%s.json" % cfg_name)
    if not os.path.exists(path):
        return None, None, None
    c = json.load(open(path))
    n = cfg["W"] * cfg["H"]
    # rays whose result cannot reach the output (tools/oracle_counters.py: "dead product" rays of the spatial pass) are
    # not cast by the product and are not charged
    trav = c["per_sample"].get("B_alg_traversal_bytes_cast", c["per_sample"]["B_alg_traversal_bytes"]) * n * spp
    per_px = 872 + 1040 * (spp - 1) + 280 * spp
    build = 424 * c["triangles"]
    if n_foreground is None:
        n_foreground = int(round(c.get("hit_fraction", 1.0) * n))
    stages = (9 + 2) * spp + (spp - 1)  # 9 forward stage entries (+ temporal for i > 0) + 2 backward per iteration
    fg = n_foreground * per_px + (n - n_foreground) * 4 * stages
    trav_ref = c["per_sample"]["B_alg_traversal_bytes"] * n * spp
    return trav + n * per_px + build, trav + fg + build, trav_ref + fg + build


def _timeline(torch, step, path):
    """Diagnostics (not a bench number): one warm step under torch.profiler (CUPTI); writes every device kernel with its
    start and duration, plus the idle gaps between kernels, so launch-bound stretches of the step can be seen."""
    from torch.profiler import profile, ProfilerActivity
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step(False)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    rows, busy, last_end = [], 0.0, t0
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        gap = max(0.0, e.time_range.start - last_end)
        last_end = max(last_end, e.time_range.end)
        busy += d
        rows.append((s, d, gap, e.name))
    with open(path, "w") as f:
        f.write("# start_us dur_us gap_before_us name ; span %.1f us, busy %.1f us, kernels %d\n" % (last_end - t0, busy, len(rows)))
        for s, d, gap, name in rows:
            f.write("%10.1f %9.1f %8.1f  %s\n" % (s, d, gap, name[:110]))


def roofline(cfg_name, cfg, per_kernel, steps, peak, peak_kind, n_foreground=None, probe_spp=None):
    """Dominant kernel vs the HBM roofline.  Algorithmic bytes per launch = S_screen(stage) * N + 36 * V_n + 48 * V_t
     with V_n / V_t the oracle's node-pop / triangle-test counts per launch under the contract schedule
    (profiles/oracle_counters_<cfg>.json, produced by tools/oracle_counters.py)."""
    if not per_kernel:
        return None
    name, (ms, count) = max(per_kernel.items(), key=lambda kv: kv[1][0])
    n = cfg["W"] * cfg["H"]
    path = os.path.join(ROOT, "profiles", "oracle_counters_%s.json" % cfg_name)
    counters = json.load(open(path)) if os.path.exists(path) else {}
    screen = {"spatial_resampling": 104, "initial_resampling": 80, "bounce_shade": 176, "bounce_first": 140,
              "final_visibility": 28, "temporal_resampling": 168}.get(name, 0)
    c = counters.get(name, {})
    if n_foreground is not None:  # per-pixel streams on foreground pixels, the occupancy word on the others
        alg = screen * n_foreground + 4 * (n - n_foreground)
    else:
        alg = screen * n
    screen_bytes = alg
    # node pops and triangle tests of the rays the product CASTS (the oracle's count without the reference's rays whose
    # result cannot reach the output, see tools/oracle_counters.py); the same figure over all the reference's rays is
    # reported beside it as frac_reference_rays
    alg += 36 * c.get("cast_nodes_per_launch", c.get("nodes_per_launch", 0)) + 48 * c.get("cast_tris_per_launch", c.get("tris_per_launch", 0))
    alg_ref = screen_bytes + 36 * c.get("nodes_per_launch", 0) + 48 * c.get("tris_per_launch", 0)
    dur = ms / count * 1e-3
    achieved = alg / dur / 1e9 if alg else None
    # DRAM traffic of the dominant kernel: dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of
    # this command, written by tools/final_round.sh together with the commit it was taken at
    traffic, traffic_source = c.get("dram_bytes_per_launch"), c.get("dram_bytes_source")
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic_%s.json" % cfg_name)
    if os.path.exists(tpath):
        t = json.load(open(tpath))
        if t.get("entry_point") == name:
            traffic, traffic_source = t.get("dram_bytes_per_launch"), "%s (commit %s)" % (t.get("source"), t.get("commit"))
    return {"kernel": name, "bound": "hbm", "achieved": achieved, "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "traffic_source": traffic_source,
            "launch_ms": ms / count, "launches_timed": count, "alg_bytes_per_launch": alg,
            "frac_reference_rays": alg_ref / dur / 1e9 / peak if alg_ref else None,
            "accounting": "achieved = S_screen * foreground pixels + 36 B per node pop + 48 B per triangle test of the rays this "
                          "kernel CASTS (oracle count, profiles/oracle_counters_<cfg>.json: cast_*), per launch, / launch_ms; "
                          "frac_reference_rays charges every ray the reference casts, including the ones the product proves "
                          "dead or already answered and skips",
            "share_of_step": ms / steps / max(sum(v[0] for v in per_kernel.values()) / steps, 1e-9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mirres_b200")
    ap.add_argument("--config", default=None, help="C2 (training step, the default), C3 / C5 (renders)")
    ap.add_argument("--mode", default=None, choices=["train", "render"],
                    help="train: stage-1 step, one view per rank (weak scaling); render: one frame cut into row bands over "
                         "the ranks (strong scaling).  Default: train for C2, render for C3 / C5")
    ap.add_argument("--spp", type=int, default=None, help="render mode: override the configuration's spp")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--emulate", default=None, help="render mode diagnostics: R/W = time the band of rank R of W on one GPU, no communication")
    ap.add_argument("--emulate-bounds", default=None, help="with --emulate: comma-separated band boundaries instead of the computed ones")
    ap.add_argument("--rebalance", type=int, default=2, help="render mode, N > 1: rounds of re-cutting the bands by measured cost")
    ap.add_argument("--uniform-bands", action="store_true", help="render mode: bands of equal height instead of equal foreground")
    ap.add_argument("--no-overlap", action="store_true", help="direct and indirect chains on one stream")
    ap.add_argument("--allreduce-after", action="store_true",
                    help="one all-reduce after the step instead of two inside it (envmap segment beside the vertex scatters)")
    ap.add_argument("--mesh-normals", action="store_true",
                    help="G-buffer normals through the reference's chain (auto_normals -> interpolation -> "
                         "prepare_shading_normal); the vertex segment of the gradient buffer then holds d loss / d vertex "
                         "positions instead of normal gradients accumulated at the vertices")
    ap.add_argument("--primary-blocks", type=int, default=4, help="blocks per SM of the primary-ray tracer (0 = library default)")
    ap.add_argument("--tune", action="append", help="library tuning NAME=VALUE (mirres_set_tuning), repeatable")
    ap.add_argument("--timeline", default=None, help="diagnostics: write a per-kernel device timeline of one warm step")
    args = ap.parse_args()
    if args.mode is None:
        args.mode = "render" if args.config in ("C3", "C5") else "train"
    if args.config is None:
        args.config = "C3" if args.mode == "render" else "C2"
    if args.impl == "reference":
        run_reference(args)
    elif args.mode == "render":
        run_render(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
