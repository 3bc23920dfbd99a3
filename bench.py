#!/usr/bin/env python3
"""bench.py — the judged benchmark of the raycast hot path (contract: see the task brief).

Headline workload (BASELINE.json configs[1]): the `xor` procedural volume, 256^3 uint8, rendered at 1920x1080
over an orbit camera sweep of 360 frames (yaw_i = 1 + 2*pi*i/360, pitch -0.5, zoom 3, the xor example's camera,
examples/xor/main.rs:273-279) in mode M1 (scalar volume, trilinear, `vertigo` transfer function, early ray
termination, exact empty-space skipping), LAYOUT_QUAD (two tex3D point fetches per sample, fp32 weights: parity path).
One STEP = one frame of the orbit; a LAUNCH renders --batch (default: up to 32) consecutive frames of the sweep (grid.z =
frame), because one 1080p frame with a fifth of its pixels on the box cannot fill a B200 (the one-frame-per-launch
figure is reported beside it as `single_frame_per_launch`). `value` = frames/s with the volume resident in HBM,
timed per launch with CUDA events on the launching stream, L2 flushed (a 256 MiB write) between timed launches.
`e2e` = the same metric through the C ABI with HOST buffers (vkrt_frames_host: cameras in, presented RGBA8 frames
out into page-locked host memory of ONE consumer process).

Beside the headline, in the same run and the same JSON line:
  roofline        the dominant kernel against the texel-path peak measured by build/microbench in this job
  roofline_dense  the same kernel with skipping off (the case where the fetch path IS the bound)
  bonsai_standin  BASELINE's metric names "bonsai 256^3 @1080p": the seeded stand-in volume beside the xor pattern
  m0_reference_exact  mode M0 = raycast_compute.wgsl literally, with the reference's own WGSL on the host cores
  configs         BASELINE configs[2..4] at THIS N: 1024^3 fp16 and 2048^3 u8 at 4K (sort-first image tiles gathered
                  to rank 0), 4096^3 fp32 sort-last (N >= 2) with the phases of a frame timed separately
  parity_checks   N >= 2: sort-first frames == single-GPU frames bit for bit, sort-last within 2/255, 0 wait timeouts

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--only-headline]

--impl reference times the CPU restatement of the reference (the oracle port; the reference itself
needs Rust + wgpu + Vulkan, none of which exist here) on all host cores, on the same workload.
N > 1 (torchrun): sort-first — groups of consecutive frames are dealt round-robin to the ranks (volume replicated),
every frame lands in rank 0's ring of frames over NVLink.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

W, H = 1920, 1080
NVOL = 256
ORBIT = 360
METRIC = "frames/s (xor 256^3 u8 @1920x1080, orbit sweep)"
WORKLOAD = "configs[1]: xor procedural volume 256^3 uint8 at 1920x1080, orbit camera sweep of 360 frames, mode M1 (trilinear + vertigo TF + ERT + exact empty-space skipping)"


def orbit_camera(rt_or_oracle, i: int):
    yaw = 1.0 + 2.0 * math.pi * (i % ORBIT) / ORBIT
    return (3.0, -0.5, yaw, (0.0, 0.0, 0.0), W / H)


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (profiling guide's clocks line)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def load_peaks() -> dict:
    peaks = {"hbm_gbs": 6650.0, "hbm_source": "fallback (B200_PROFILING.md)"}
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            peaks["hbm_gbs"] = float(json.loads(p.read_text())["hbm_gbs"])
            peaks["hbm_source"] = "MEASURED_PEAKS.json"
        except Exception:
            pass
    m = ROOT / "profiles" / "microbench_r01.json"
    if m.exists():
        try:
            peaks["micro"] = json.loads(m.read_text())
            peaks["micro_source"] = "profiles/microbench_r01.json (committed copy, measured on this pool's B200 in round 1)"
        except Exception:
            pass
    return peaks


def run_microbench(device: int):
    """The texel-path / L1 / L2 / HBM peaks of THIS GPU, measured in THIS job by build/microbench (bench/microbench.cu),
    outside every timed region. Returns (dict, clocks) or (None, None)."""
    exe = ROOT / "build" / "microbench"
    if not exe.exists():
        return None, None
    clocks = ClockSampler(device)
    clocks.start()
    try:
        r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=240)  # device 0 = the N=1 bench device
        out = json.loads(r.stdout) if r.returncode == 0 else None
    except Exception:
        out = None
    return out, clocks.stop()


# ------------------------------------------------------------------------------------------------
def host_cores() -> int:
    """Host cores this process may use — NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_run(steps: int, warmup: int, tiles_per_step: int = 4) -> dict:
    """The reference's CPU arm: oracle port of the M1 march on all host cores. Each step renders a
    stratified 1/36 sample of one orbit frame (4 of 144 tiles of 160x120 px... see `sample`)."""
    from oracle import binding as ob
    from vokselis_b200 import abi, volumes

    vol = volumes.xor_u8(NVOL)
    p = abi.default_params(abi.MODE_M1)
    ts = 120
    p.tile_size = ts
    cols, rows = W // ts, H // ts  # 16 x 9 = 144 tiles cover 1920x1080 exactly
    ntiles = cols * rows
    stride = ntiles // tiles_per_step
    frame = np.zeros((H, W, 4), np.uint16)
    cores = host_cores()  # passed to the oracle explicitly (num_threads(nt)), whatever OMP_NUM_THREADS says

    def step(i):
        cam = ob.camera_uniform(*orbit_camera(None, i))
        ids = [(i * 7 + k * stride) % ntiles for k in range(tiles_per_step)]
        offs = np.array([[(t % cols) * ts, (t // cols) * ts] for t in ids], np.float32)
        _, _, st = ob.render(p, cam, W, H, scalar=vol, offsets=offs, frame=frame, want_aux=False, nthreads=cores)
        return st.samples_reference

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    samples = 0
    for i in range(steps):
        samples += step(warmup + i)
    dt = time.perf_counter() - t0
    frames_equiv = steps * tiles_per_step / ntiles
    return {"fps": frames_equiv / dt, "seconds": dt, "samples_per_s": samples / dt, "cores": cores,
            "sample": f"{steps} steps x {tiles_per_step} of {ntiles} tiles ({ts}x{ts} px, stratified, rotating with the orbit) of the 1920x1080 frame "
                      f"= {frames_equiv:.2f} frame-equivalents in {dt:.1f} s"}


def m0_section(rt, abi, device, K, Wm, no_cpu):
    """Mode M0 — shaders/raycast_compute.wgsl literally (two rgba16f 256^3 volumes from the xor generator,
    nearest fetch, dt floor 0.01): device-timed frames/s at 1920x1080 and at the reference's own 1280x720,
    L2 flushed between frames, and beside it the reference's OWN WGSL (machine-translated, oracle/_ref) on
    the host cores at 1280x720 — the closest available stand-in for 'wgpu on lavapipe'."""
    out = {}
    color = normal = None
    for (w, h) in ((1920, 1080), (1280, 720)):
        c = rt.Context(device, w, h)
        c.generate_xor(256, 0)
        q = rt.default_params(abi.MODE_M0)
        q.skip_empty, q.layout = 1, abi.LAYOUT_TEXTURE
        c.set_params(q)
        cams = [rt.Camera(3.0, -0.5, 1.0 + 2.0 * math.pi * i / ORBIT, (0.0, 0.0, 0.0), w / h).get_proj_view_matrix() for i in range(ORBIT)]
        n = min(K, 120)
        c.timing_enable(n)
        for i in range(min(Wm, 10)):
            c.flush_l2()
            c.render(cams[i])
        for i in range(n):
            c.flush_l2()
            c.render(cams[(Wm + i) % ORBIT])
        ms = c.timing_read(n).astype(np.float64)
        out[f"gpu_fps_{w}x{h}_one_frame_per_launch"] = 1e3 / float(ms.mean())
        out[f"gpu_ms_{w}x{h}_one_frame_per_launch"] = float(ms.mean())
        # MAX_BATCH frames of the sweep per launch (grid.z = frame), like the headline
        FB = rt.MAX_BATCH
        nb = max(n // FB, 2)
        c.timing_enable(nb)
        c.render_batch(cams[:FB])
        for j in range(nb):
            c.flush_l2()
            c.render_batch([cams[(Wm + FB * j + k) % ORBIT] for k in range(FB)])
        msb = c.timing_read(nb).astype(np.float64)
        out[f"gpu_fps_{w}x{h}"] = FB * 1e3 / float(msb.mean())
        out[f"gpu_ms_{w}x{h}"] = float(msb.mean()) / FB
        if (w, h) == (1280, 720) and not no_cpu:
            color, normal = c.download_rgba16f()
            cam0 = cams[0]
        c.close()
    out["layout"] = f"TEXTURE (tex3D point fetches), exact empty-space skipping; gpu_fps_* = {rt.MAX_BATCH} frames per launch, L2 flushed between launches"
    if color is not None:
        try:
            from oracle import ref_binding as rb

            if rb.available():
                nt = host_cores()
                rb.raycast_compute(cam0, color, normal, 1280, 720, nthreads=nt)  # warm-up
                t0 = time.perf_counter()
                reps = 2
                for _ in range(reps):
                    rb.raycast_compute(cam0, color, normal, 1280, 720, nthreads=nt)
                dt = (time.perf_counter() - t0) / reps
                out["cpu_reference_wgsl"] = {"value": 1.0 / dt, "unit": "frames/s", "cores": nt, "kind": "reference",
                                             "sample": f"{reps} full 1280x720 frames of `single` (xor camera), oracle/_ref = the reference's "
                                                       "raycast_compute.wgsl machine-translated to C++ (oracle/wgsl2cpp.py), OpenMP over rows"}
            else:
                out["cpu_reference_wgsl"] = {"unavailable": "oracle/_ref not built (needs /root/reference at build time)"}
        except Exception as e:  # the checker must never break the bench
            out["cpu_reference_wgsl"] = {"unavailable": repr(e)}
    return out


def host_sweep(ctx, rt, cams, frame_ids, per_call, group):
    """Frames `frame_ids` of the orbit through vkrt_frames_host into page-locked host memory, `per_call` frames per
    blocking call; returns the summed wall time of the calls (the L2 flush before each call is not timed)."""
    if not frame_ids:
        return 0.0
    pinned = rt.PinnedArray((per_call, H, W, 4), np.uint8)
    out = pinned.array
    ctx.frames_host([cams[i % ORBIT] for i in range(per_call)], out, group=group)  # warm-up: buffers, clocks
    tot, done = 0.0, 0
    while done < len(frame_ids):
        n = min(per_call, len(frame_ids) - done)
        cs = [cams[f] for f in frame_ids[done:done + n]]
        ctx.flush_l2()
        ctx.sync()
        t0 = time.perf_counter()
        ctx.frames_host(cs, out[:n], group=group)  # blocking: returns when the n frames are in host memory
        tot += time.perf_counter() - t0
        done += n
    pinned.close()
    return tot


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # ~10 frame-equivalents in total whatever K is (a 1080p M1 frame costs the oracle ~0.3-0.6 s on 16-32 cores)
    tiles = max(4, min(144, int(round(1440 / max(args.steps, 1)))))
    r = cpu_reference_run(args.steps, args.warmup, tiles_per_step=tiles)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["fps"], "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * r["seconds"] / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "arm": "CPU restatement of the reference shader (oracle port, OpenMP) — the Rust/wgpu reference "
                   "cannot be built in this image; substitute for wgpu-on-lavapipe", "step": "bounded sample of one frame, see cpu_baseline.sample"},
        "ray_samples_per_s": r["samples_per_s"],
        "cpu_baseline": {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["fps"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------------
def device_timed_batches(ctx, cams, first, n_launches, B, params, flush=True):
    """n_launches launches of B consecutive orbit frames each (vkrt_render_batch), CUDA events per launch on the context's
    stream, L2 flushed before each (untimed). Returns the per-launch ms."""
    ctx.set_params(params)
    ctx.render_batch([cams[(first + k) % ORBIT] for k in range(B)])  # warm-up: layout build, module load
    ctx.render_batch([cams[(first + k) % ORBIT] for k in range(B)])
    ctx.timing_enable(n_launches)
    for j in range(n_launches):
        if flush:
            ctx.flush_l2()
        ctx.render_batch([cams[(first + j * B + k) % ORBIT] for k in range(B)])
    return ctx.timing_read(n_launches).astype(np.float64)


def probe_samples(ctx, rt, abi, cams, frame_ids, layout, skip):
    """Mean reference-semantics and fetched samples per frame over `frame_ids` (counting kernel, untimed)."""
    q = rt.default_params(abi.MODE_M1)
    q.skip_empty, q.count_samples, q.layout = skip, 1, layout
    ctx.set_params(q)
    tot_ref = tot_fetched = 0
    for f in frame_ids:
        ctx.reset_stats()
        ctx.render(cams[f % ORBIT])
        st = ctx.stats()
        tot_ref += st.samples_reference
        tot_fetched += st.samples_fetched
    return tot_ref / len(frame_ids), tot_fetched / len(frame_ids)


def run_gpu(args):
    import torch
    import torch.distributed as dist

    from vokselis_b200 import abi, rt, volumes, workloads

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at first use; stdout must carry exactly one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    rt.lib()  # fail loudly if the CUDA library is missing
    K, Wm = args.steps, args.warmup
    peaks = load_peaks()
    LAYOUT = abi.LAYOUT_QUAD

    vol = volumes.xor_u8(NVOL)
    cams = [rt.Camera(*orbit_camera(rt, i)).get_proj_view_matrix() for i in range(ORBIT)]
    ctx = rt.Context(local, W, H)
    ctx.upload_scalar(vol)
    p = rt.default_params(abi.MODE_M1)
    p.skip_empty = 1
    p.layout = LAYOUT  # exact fp32-weight trilinear from two tex3D point fetches of pre-gathered xy quads (parity-grade)
    ctx.set_params(p)

    def barrier():
        ctx.sync()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    # ---- sample statistics of the workload (untimed, counting kernel) over exactly the cameras the timed passes render ----
    samples_ref = samples_fetched = 0
    timed_ids = [(Wm + i) % ORBIT for i in range(K)]
    if rank == 0:
        samples_ref, samples_fetched = probe_samples(ctx, rt, abi, cams, timed_ids, LAYOUT, 1)
        ctx.set_params(p)

    # frames per launch (grid.z = frame). --batch 0 = choose: the fewest launches of <= MAX_BATCH = 32 frames on one GPU (8 / 16 / 24 / 32 frames per launch: 10,065 / 10,680 / 11,086 / 11,090 frames/s); for N ranks the group size that leaves no rank
    # with more frames than necessary (groups are dealt round-robin), larger groups preferred
    if args.batch > 0:
        B = max(1, min(args.batch, rt.MAX_BATCH))
    elif world == 1:
        B = -(-K // -(-K // rt.MAX_BATCH))  # the fewest launches of <= MAX_BATCH frames, evenly filled (20 steps: 2 launches of 10)
    else:
        from vokselis_b200.sortfirst import choose_batch

        B = choose_batch(K, world)
    L = (K + B - 1) // B                        # launches per timed pass

    def chunk(i0):  # cameras of the launch that starts at step i0 (the orbit wraps)
        return [cams[(Wm + i0 + k) % ORBIT] for k in range(min(B, K - i0))]

    if world > 1:
        from vokselis_b200 import sortfirst

        # 1080p frames take a fraction of a millisecond on one GPU: deal whole frames — groups of B consecutive
        # frames, one launch per group — round-robin; every frame still lands in rank 0's ring (peers ship a group
        # with one copy-engine transfer over NVLink, overlapping their next launch).
        group = sortfirst.SortFirstGroup(ctx, rank, world, granularity=args.granularity, tile=120, batch=B if args.granularity == "frames" else 1)
        GB = group.batch

        def launch(i0, flush):
            if group.granularity == "tiles":
                group.render(cams[(Wm + i0) % ORBIT])
            else:  # the owner flushes on the stream its launch uses (the root renders on its second stream)
                group.render_batch(chunk(i0) if GB > 1 else [cams[(Wm + i0) % ORBIT]], flush_l2=flush)
        step_stride = GB if group.granularity == "frames" else 1
    else:
        group = None

        def launch(i0, flush):
            if flush:
                ctx.flush_l2()
            ctx.render_batch(chunk(i0))
        step_stride = B

    # ---- timed: device time per launch (CUDA events on the launching stream), L2 flushed between launches -----
    ctx.timing_enable(max(K, 1))
    # warm-up: at least Wm steps (>= 3), and with N ranks at least two launches on EVERY rank (buffers, module load, clocks)
    n_warm = max(-(-max(Wm, 3) // step_stride), 1)
    if group is not None and group.granularity == "frames":
        n_warm = max(n_warm, 2 * world)
    for j in range(n_warm):
        launch((j * step_stride) % max(K - step_stride + 1, 1), True)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()

    def timed_pass(flush):
        """K steps (frames) between barriers; returns (device ms of the launches THIS rank issued, wall s, ms on rank 0's
        own stream from the first wait to the last consume — the frames-gathered-on-rank-0 timeline at N > 1)."""
        first = group.frame if group is not None else 0
        barrier()
        t0 = time.perf_counter()
        ctx.mark(0)
        for i0 in range(0, K, step_stride):
            launch(i0, flush)
        ctx.mark(1)
        barrier()
        wall = time.perf_counter() - t0
        timeline = ctx.mark_elapsed(0, 1)
        if group is None:
            mine = len(range(0, K, step_stride))
        elif group.granularity == "tiles":
            mine = K
        else:
            mine = group.my_launches(first, group.frame - first)
        return (ctx.timing_read(mine).astype(np.float64) if mine > 0 else np.zeros(0)), wall, timeline

    def whole_job_ms(ms):
        total = float(ms.sum())  # this rank's busy device time: its launches (the group transfers overlap them on the copy stream)
        if world > 1:
            t = torch.tensor([total], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t.item())
        return total

    launch_ms, t_wall, timeline_ms = timed_pass(True)
    total_ms = whole_job_ms(launch_ms)
    fps = K / (total_ms * 1e-3)
    # warm-L2 variant (steady-state orbit, no flush), device-timed the same way
    warm_ms, t_wall_warm, timeline_warm_ms = timed_pass(False)
    warm_total_ms = whole_job_ms(warm_ms)

    # ---- one frame per launch (latency of a single frame), N = 1 only ---------------------------------------
    single = None
    if world == 1:
        n1 = min(K, 120)
        ctx.timing_enable(n1)
        for i in range(n1):
            ctx.flush_l2()
            ctx.render(cams[(Wm + i) % ORBIT])
        s_ms = ctx.timing_read(n1).astype(np.float64)
        single = {"ms_per_frame": float(s_ms.mean()), "frames_per_s": 1e3 / float(s_ms.mean()),
                  "p10_p50_p90_ms": [float(np.percentile(s_ms, q)) for q in (10, 50, 90)],
                  "note": "vkrt_render, one frame per launch, L2 flushed between frames: a 1080p frame with a fifth of its pixels on the box "
                          "does not fill 148 SMs and is bounded by the dependent march of its longest rays"}

    # ---- e2e through the C ABI with host buffers: presented RGBA8 frames in ONE consumer's page-locked host memory ----------
    e2e = None
    PER_CALL = 60  # frames per blocking vkrt_frames_host call: a 360-frame sweep in 6 calls (20 driver steps: one call)
    if world == 1:
        # the call a user of a sweep makes: vkrt_frames_host — cameras in, presented RGBA8 frames out in page-locked host memory;
        # groups of frames per launch, present fused into the raycast epilogue, D2H of a group overlapping the next raycast
        tot = host_sweep(ctx, rt, cams, [(Wm + i) % ORBIT for i in range(K)], PER_CALL, min(B, 4))
        e2e_frames = K / tot
        # one frame per call (vkrt_frame_host, blocking), for comparison
        pinned1 = rt.PinnedArray((H, W, 4), np.uint8)
        n1 = min(K, 120)
        for i in range(3):
            ctx.frame_host(cams[i], pinned1.array)
        tot1 = 0.0
        for i in range(n1):
            ctx.flush_l2()
            ctx.sync()
            t0 = time.perf_counter()
            ctx.frame_host(cams[(Wm + i) % ORBIT], pinned1.array)
            tot1 += time.perf_counter() - t0
        # the PCIe floor, measured here: the same bytes D2H into page-locked memory with no rendering
        nfl = min(K, 96)
        pinned_fl = rt.PinnedArray((8, H, W, 4), np.uint8)
        ctx.sync()
        t0 = time.perf_counter()
        for i in range(nfl):
            ctx.readback_rgba8_async(pinned_fl.array[i % 8])
        ctx.sync()
        d2h_floor_fps = nfl / (time.perf_counter() - t0)
        pinned_fl.close()
        pinned1.close()
        e2e = {"value": e2e_frames, "unit": "frames/s", "h2d_bytes_per_step": 144 + 48, "d2h_bytes_per_step": W * H * 4,
               "how": f"vkrt_frames_host, {PER_CALL} frames per blocking call into page-locked host memory: cameras in (kernel arguments), "
                      f"groups of {min(B, 4)} frames per launch with the present pass fused into the raycast epilogue, RGBA8 D2H of a group overlapping "
                      "the raycast of the next; wall clock per call, L2 flushed before each call (flush untimed)",
               "single_frame_blocking": n1 / tot1,
               "d2h_floor": {"frames_per_s": d2h_floor_fps, "GBs": d2h_floor_fps * W * H * 4 / 1e9, "of_floor": e2e_frames / d2h_floor_fps,
                             "note": "measured in this run: the same RGBA8 bytes copied D2H into page-locked memory back to back, no rendering"},
               "note": "single_frame_blocking = vkrt_frame_host, one frame per call (raycast + fused present + D2H, nothing overlapped). "
                       "The PCIe floor for 8.3 MB of RGBA8 per frame is ~0.154 ms (54 GB/s measured) = 6,500 frames/s per link"}
    else:
        funnel = group.e2e(cams, K, Wm)   # every frame through rank 0's GPU and its single PCIe link (kept for comparison)
        timeouts = ctx.sortfirst_timeouts() if rank == 0 else 0
        group.close()
        # The gather a host consumer wants: ONE page-locked host segment owned by rank 0's process, mapped and registered by
        # every rank (POSIX shared memory + cudaHostRegister); every rank renders a contiguous 1/N of the sweep with
        # vkrt_frames_host straight into it over its OWN PCIe link. Same meaning as at N = 1: presented RGBA8 frames in the
        # consumer's host memory.
        n_seg = min(K, 96)  # ring of host frames (frame i of the sweep -> slot i % n_seg)
        seg = rt.SharedHostFrames(rank, world, dist, n_seg, H, W)
        lo_f, hi_f = rank * K // world, (rank + 1) * K // world
        ids = list(range(lo_f, hi_f))
        my_tot = 0.0
        if ids:
            ctx.frames_host([cams[(Wm + i) % ORBIT] for i in ids[:min(4, len(ids))]], seg.array[:min(4, len(ids))], group=4)  # warm-up
        barrier()
        done = 0
        while done < len(ids):
            s0 = ids[done] % n_seg
            n = min(PER_CALL, len(ids) - done, n_seg - s0)  # contiguous run of ring slots
            cs = [cams[(Wm + i) % ORBIT] for i in ids[done:done + n]]
            ctx.flush_l2()
            ctx.sync()
            t0 = time.perf_counter()
            ctx.frames_host(cs, seg.array[s0:s0 + n], group=4)
            my_tot += time.perf_counter() - t0
            done += n
        tt = torch.tensor([my_tot], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        barrier()
        # the consumer's check (untimed): the LAST frame of every rank's share, read from the shared segment by rank 0,
        # equals rank 0's own rendering of that camera
        seg_ok = None
        if rank == 0:
            seg_ok = True
            for r in range(world):
                last = (r + 1) * K // world - 1
                if last < r * K // world:
                    continue
                mine8 = ctx.frames_host([cams[(Wm + last) % ORBIT]], group=1)[0]
                seg_ok = seg_ok and bool(np.array_equal(mine8, seg.array[last % n_seg]))
        # What the platform allows: the SAME bytes into the SAME segment with no rendering at all — every rank copies its share of
        # the sweep (its presented frame, again and again) over its own PCIe link at the same time, 3 passes, slowest rank.
        floor_tot = 0.0
        barrier()  # rank 0 has finished reading the rendered frames: the copies below overwrite them
        if ids:
            ctx.frame_host(cams[Wm % ORBIT], seg.array[ids[0] % n_seg])  # something presented in ctx's rgba8 buffer
        for _ in range(3):
            barrier()
            t0 = time.perf_counter()
            for i in ids:
                ctx.readback_rgba8_async(seg.array[i % n_seg])
            ctx.sync()
            floor_tot += time.perf_counter() - t0
        ft = torch.tensor([floor_tot / 3.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(ft, op=dist.ReduceOp.MAX)
        barrier()
        d2h_floor_fps = K / float(ft.item()) if float(ft.item()) > 0 else None
        seg.close()
        e2e = {"value": K / float(tt.item()) if float(tt.item()) > 0 else 0.0, "unit": "frames/s", "h2d_bytes_per_step": 144 + 48,
               "d2h_bytes_per_step": W * H * 4,
               "how": f"presented RGBA8 frames gathered in ONE page-locked host segment owned by rank 0's process (POSIX shared memory, cudaHostRegister'ed "
                      f"in every rank): every rank renders a contiguous 1/{world} of the sweep with vkrt_frames_host ({PER_CALL} frames per blocking call, groups of 4 "
                      "frames per launch, present fused, D2H overlapping the next group) straight into it over its own PCIe link; K frames over the slowest rank's "
                      "summed call time, L2 flushed before each call (flush untimed)",
               "consumer_sees_every_ranks_frames": seg_ok, "host_ring_frames": n_seg,
               "d2h_floor": {"frames_per_s": d2h_floor_fps, "GBs": (d2h_floor_fps * W * H * 4 / 1e9) if d2h_floor_fps else None,
                             "of_floor": (K / float(tt.item()) / d2h_floor_fps) if d2h_floor_fps and float(tt.item()) > 0 else None,
                             "note": f"measured in this run: the {world} ranks copy the same {K} frames' bytes into the same host segment concurrently, no rendering "
                                     "(host memory / PCIe root complexes are shared by the GPUs: the links do not add up)"},
               "via_rank0_gpu": funnel}

    clock_info = clocks.stop() if rank == 0 else None  # sampled across all timed GPU regions above (value, warm, e2e)

    # ---- beside the headline, N = 1 only: dense case, bonsai stand-in, reference-exact mode, CPU baseline, this GPU's peaks ------
    cpu = m0 = dense = bonsai = micro = micro_clocks = None
    if world == 1:
        nb = max(min(L, 12), 3)
        pd = rt.default_params(abi.MODE_M1)
        pd.skip_empty, pd.layout = 0, LAYOUT
        d_ms = device_timed_batches(ctx, cams, Wm, nb, B, pd)
        d_ref, d_fetched = probe_samples(ctx, rt, abi, cams, [(Wm + i) % ORBIT for i in range(nb * B)], LAYOUT, 0)
        dense = {"ms_per_frame": float(d_ms.sum()) / (nb * B), "samples_per_frame": d_fetched, "launches": nb, "frames_per_launch": B}
        # bonsai stand-in (the dataset is absent from the reference: .MISSING_LARGE_BLOBS), same camera sweep
        ctx.upload_scalar(volumes.bonsai_standin_u8(NVOL, seed=1))
        bonsai = {"volume": "vokselis_b200.volumes.bonsai_standin_u8(256, seed=1): stand-in for bonsai_256x256x256_uint8.raw, which the reference does not ship"}
        for skip in (1, 0):
            pb = rt.default_params(abi.MODE_M1)
            pb.skip_empty, pb.layout = skip, LAYOUT
            b_ms = device_timed_batches(ctx, cams, Wm, nb, B, pb)
            b_ref, b_fetched = probe_samples(ctx, rt, abi, cams, [(Wm + i) % ORBIT for i in range(0, nb * B, 4)], LAYOUT, skip)
            msf = float(b_ms.sum()) / (nb * B)
            bonsai["skip" if skip else "no_skip"] = {"frames_per_s": 1e3 / msf, "ms_per_frame": msf, "samples_per_frame": {"reference": b_ref, "fetched": b_fetched},
                                                    "ray_samples_per_s": b_ref * 1e3 / msf, "fetches_per_s": 2.0 * b_fetched * 1e3 / msf}
        ctx.upload_scalar(vol)
        ctx.set_params(p)
        m0 = m0_section(rt, abi, local, K, Wm, args.no_cpu)
        micro, micro_clocks = run_microbench(local)
        if not args.no_cpu:
            r = cpu_reference_run(steps=36, warmup=2, tiles_per_step=8)
            cpu = {"value": r["fps"], "unit": "frames/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
                   "ray_samples_per_s": r["samples_per_s"],
                   "note": "CPU restatement of the reference shader (oracle port) — substitute for wgpu/lavapipe, which cannot be installed here"}
    ctx.close()

    # ---- BASELINE configs[2..4] at this N, and the multi-GPU correctness checks, in the same run ------------------------------
    configs, parity = {}, {}
    if not args.only_headline:
        if world > 1:
            for name, fn in (("sortfirst", workloads.check_sortfirst), ("sortlast", workloads.check_sortlast)):
                try:
                    parity[name] = fn(rank, world, local, dist)
                except Exception as e:  # a failed check must show up in the line, not kill it
                    parity[name] = {"ok": False, "error": repr(e)}
        else:
            parity["note"] = "sort-first / sort-last checks need >= 2 GPUs; single-GPU parity against the oracle is the -m gpu test tier"
        fr = max(8, min(args.config_frames, 48))
        for cid, key in ((3, "config3_1024_f16_4k_sortfirst_tiles"), (4, "config4_2048_u8_sparse_4k_sortfirst_tiles")):
            try:
                r = workloads.run_sortfirst_tiles(cid, rank, world, local, dist if world > 1 else None, frames=fr, hbm_peak_gbs=peaks["hbm_gbs"])
            except Exception as e:
                r = {"error": repr(e)}
            configs[key] = r
        if world > 1:
            try:
                configs["config5_4096_f32_4k_sortlast"] = workloads.run_sortlast(rank, world, local, dist, edge=4096, frames=6, hbm_peak_gbs=peaks["hbm_gbs"])
            except Exception as e:
                configs["config5_4096_f32_4k_sortlast"] = {"error": repr(e)}
        else:
            configs["config5_4096_f32_4k_sortlast"] = {"not_run": "256 GiB exceeds one GPU's 180 GB: sort-last needs N >= 2 (see the N = 2/4/8 lines)"}

    if rank == 0:
        ms = total_ms / K
        frames_per_launch = step_stride
        # Roofline of the dominant kernel, raycast_kernel<M1, QUAD, SKIP>.
        # Algorithmic bytes per ray-sample: 8 taps x 1 B; units per launch = the samples the kernel actually
        # fetches for its frames. The 16 MiB volume (64 MiB of pre-gathered quads) is L1/L2-resident, so the texture path
        # (two point fetches of 4-byte texels per sample) is the binding memory resource, not HBM; both are reported.
        # achieved = algorithmic bytes of the frames ACTUALLY launched / their summed launch time (rank 0's launches; a last,
        # shorter launch counts with its own frames). At N = 1 this equals samples_per_frame.fetched x 8 B / ms_per_step.
        my_frames = K if world == 1 else max(int(round(K * len(launch_ms) / max(L, 1))), 1)
        sum_ms = float(launch_ms.sum())
        kernel_ms = sum_ms / max(len(launch_ms), 1)   # average launch duration (CUDA events on the launching stream)
        micro_source = "build/microbench (bench/microbench.cu) run by this bench.py process on this GPU right after the timed regions"
        if not micro:
            micro, micro_source = peaks.get("micro", {}), peaks.get("micro_source", "none")
        alg_bytes_total = samples_fetched * 8.0 * my_frames
        alg_bytes = alg_bytes_total / max(len(launch_ms), 1)  # per (average) launch
        achieved = alg_bytes_total / (sum_ms * 1e-3) / 1e9
        fetch_peak = micro.get("tex3d_point_rgba8_F16_gfetch_s") or micro.get("tex3d_point_rgba16f_F16_gfetch_s")  # G fetches/s; 4 B of taps each
        l1_peak = 4.0 * fetch_peak if fetch_peak else None
        traffic = None
        tfile = ROOT / "profiles" / "traffic_r02.json"
        if tfile.exists():
            try:
                traffic = json.loads(tfile.read_text()).get(f"raycast_m1_quad_u8_skip_batch{frames_per_launch}_dram_bytes_per_launch")  # ncu capture of this launch shape, or null
            except Exception:
                pass
        hbm_bytes = min(4 * NVOL ** 3, alg_bytes) + (my_frames / max(len(launch_ms), 1)) * W * H * 8.0
        roofline = {
            "kernel": "raycast_kernel<M1, QUAD, SKIP>", "bound": "tex",
            "achieved": achieved, "peak": l1_peak, "unit": "GB/s", "frac": (achieved / l1_peak) if l1_peak else None, "traffic": traffic,
            "peak_source": "tex3D point fetch of rgba8 texels (coherent 8x4 footprints, L1-resident) x 4 B from %s; tld4 %s G/s, tex3D trilinear %s Gfetch/s for comparison"
                           % (micro_source, micro.get("tld4_a2d_u8_F16_ginstr_s"), micro.get("tex3d_linear_u8_F16_gfetch_s")),
            "peak_clocks": micro_clocks, "microbench": micro if world == 1 else None,
            "algorithmic_bytes_per_launch": alg_bytes, "kernel_ms": kernel_ms,
            "hbm": {"bound": "hbm", "achieved": hbm_bytes / (kernel_ms * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": hbm_bytes / (kernel_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], "peak_source": peaks["hbm_source"],
                    "compulsory_bytes_per_launch": hbm_bytes, "note": "quad texture (64 MiB, read at most once per launch) + frames (W*H*8 B each); far below HBM peak by construction"},
            "frames_per_launch": frames_per_launch,
            "binding_resource": "instruction issue: ncu on the 16-frame launch reports issue active 83 %, sm__throughput 82 % of peak over the launch, l1tex 62 %, "
                                "DRAM 2.6 %, 24.6 of 32 lanes active per instruction, 63.2 M warp instructions per frame (profiles/r02_v6_prof_batch16_m1_quad_skip_final.md); the quad "
                                "texture is L1/L2-resident, so the texel path is the memory-side bound reported here and HBM (roofline.hbm) is a few % by construction. With skipping "
                                "off the SAME kernel is bound by the texel path: roofline_dense",
            "note": "achieved = samples actually fetched x 8 B of taps / launch time. Exact empty-space skipping over 2-voxel occupancy bricks with one distance field per ray "
                    "octant fetches 11.3 M of the frame's 90 M reference samples (8-voxel bricks, isotropic field: 17.7 M): the kernel got 1.33x faster (11,215 -> 14,950 frames/s) "
                    "by NOT fetching transparent samples, which lowers this fraction (0.37 -> 0.33) while frames/s rise; what remains per frame is traversal + shading, bound by "
                    "instruction issue (profiles/r02_octant_ab.md)",
        }
        roofline_dense = None
        if dense and l1_peak:
            a = dense["samples_per_frame"] * 8.0 / (dense["ms_per_frame"] * 1e-3) / 1e9
            roofline_dense = {"kernel": "raycast_kernel<M1, QUAD, no skip>", "bound": "tex", "achieved": a, "peak": l1_peak, "unit": "GB/s", "frac": a / l1_peak,
                              "frames_per_s": 1e3 / dense["ms_per_frame"], "ms_per_frame": dense["ms_per_frame"], "samples_per_frame": dense["samples_per_frame"],
                              "fetches_per_s": 2.0 * dense["samples_per_frame"] * 1e3 / dense["ms_per_frame"],
                              "note": "the same workload with skipping off (every reference sample fetched): ncu l1tex throughput 98 %, tex_throttle the top stall "
                                      "(profiles/r02_v4_prof_batch16_m1_quad_noskip.md) — the fetch path is the bound here"}
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": Wm, "ms_per_step": ms,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "volume": "xor bit pattern (shaders/xor.wgsl:46-53) quantised to u8, 16 MiB", "resolution": [W, H],
                       "l2": "flushed between timed launches (write of a 256 MiB buffer, untimed)",
                       "layout": "QUAD (pre-gathered xy quads in a 3-D rgba8 texture, two tex3D point fetches per sample, fp32 weights: parity path)",
                       "frames_per_launch": frames_per_launch, "step": "one frame of the orbit; a launch renders frames_per_launch consecutive frames (grid.z = frame), every frame bit-identical to a single-frame launch",
                       "parallelism": "single GPU" if world == 1 else
                       f"sort-first over {world} GPUs ({args.granularity} dealt round-robin in groups of {frames_per_launch}), volume replicated, frames gathered in rank 0's ring over NVLink"},
            "ray_samples_per_s": samples_ref * fps, "fetched_samples_per_s": samples_fetched * fps,
            "samples_per_frame": {"reference": samples_ref, "fetched": samples_fetched},
            "ms_per_step_warm_l2": warm_total_ms / K, "fps_warm_l2": K / (warm_total_ms * 1e-3),
            "wall_ms_per_step_incl_flush": 1e3 * t_wall / K, "wall_ms_per_step_warm_l2": 1e3 * t_wall_warm / K,
            "rank0_timeline_ms_per_step": {"flushed": timeline_ms / K, "warm_l2": timeline_warm_ms / K,
                                           "note": "CUDA events on rank 0's own stream around the whole timed pass: at N > 1 that stream waits for and consumes every "
                                                   "frame in order, i.e. frames gathered on rank 0 per second (L2 flushes of the launching ranks included)"},
            "launch_ms_p10_p50_p90": [float(np.percentile(launch_ms, q)) for q in (10, 50, 90)] if len(launch_ms) else None,
            "single_frame_per_launch": single,
            "roofline": roofline, "roofline_dense": roofline_dense, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": len(range(0, K, step_stride)), "clocks": clock_info,
            "sortfirst_wait_timeouts": (timeouts if world > 1 else None),
            "bonsai_standin": bonsai, "m0_reference_exact": m0,
            "configs": configs, "parity_checks": parity,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=360)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg (development)")
    ap.add_argument("--granularity", default="frames", choices=["frames", "tiles"], help="sort-first granularity for N > 1")
    ap.add_argument("--batch", type=int, default=0, help="frames per launch (grid.z = frame), 1..32; 0 = choose (the fewest, evenly filled launches on one GPU)")
    ap.add_argument("--only-headline", action="store_true", help="skip BASELINE configs[2..4] and the multi-GPU checks (development)")
    ap.add_argument("--config-frames", type=int, default=24, help="frames per timed pass of configs[2]/[3]")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
