#!/usr/bin/env python
"""Benchmark of the ray-integration hot path (BASELINE.json metric: ray-steps/s and rays/s for rays
through a Gaussian-random ne cube on B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c3|c2|c1]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one bundle of rays: calc_dndr (ne cube -> float4 gradient
grid), Morton sort of the launch rays, the RK4 trace kernel, Shadowgraphy optics + histogram, and
(N > 1) the NCCL all-reduce of the histogram.  Rays are sharded per rank (weak scaling: every GPU
traces RAYS rays), the ne cube is broadcast from rank 0 once.

  value : whole-job ray-steps/s with inputs (ne cube, launch rays) resident in HBM
  e2e   : the same metric through the public Python API with HOST (pinned) numpy buffers in and the
          histogram + ray-step count read back to the host inside the timed region
  roofline : trace kernel, algorithmic bytes = 512 B per ray-step (4 stages x 8 corners x 16 B,
          SURVEY section 8d) over the kernel's CUDA-event duration, against the measured HBM copy
          bandwidth of MEASURED_PEAKS.json
  cpu_baseline : the reference's scipy path (oracle port: solve_ivp RK45 at default tolerances over
          RegularGridInterpolator, multiprocessing.Pool over ray bundles as example_multiprocess.py)
          on a bounded sample of the same cube, on this box's host cores
  cpu_baseline_c : the same sample through the plain-C restatement of that path (oracle/tt_oracle.c,
          one thread per bundle) -- the stronger CPU baseline, reported next to the scipy one

--impl reference times that CPU path alone (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (N_half of gaussian3D_FFT -> M = 2N+1 points per axis, rays per GPU, BASELINE config)
    "c3": (256, 100_000_000, "configs[2]: 513^3 (gaussian3D_FFT N=256) k^-11/3 GRF ne cube, 1e8 rays per GPU, shadowgraphy"),
    "c2": (128, 10_000_000, "configs[1]: 257^3 (gaussian3D_FFT N=128) k^-11/3 GRF ne cube, 1e7 rays per GPU, shadowgraphy"),
    "c5": (512, 125_000_000, "configs[4] per-GPU share: 1025^3 (gaussian3D_FFT N=512) k^-11/3 GRF ne cube (17.2 GB float4 grid), 1.25e8 rays per GPU (1e9 on 8), shadowgraphy"),
    "c4": (128, 10_000_000, "configs[3]: 257^3 GRF ne cube + B cube (10 T along the beam + GRF perturbation) + Te cube, 1e7 rays per GPU, "
                            "phase + Faraday rotation + inverse-bremsstrahlung attenuation, shadowgraphy"),
    "c1": (32, 100_000, "smoke-size: 65^3 GRF cube, 1e5 rays"),
}
BEAM_SIZE, DIVERGENCE, EXTENT, LWL = 4e-3, 0.05e-3, 5e-3, 1053e-9
SPECTRUM = lambda k: k ** (-11.0 / 3.0)          # noqa: E731


def ne_from_field(f, xp):
    """ne = 1e25 * clip(1 + 0.3 f / sigma, 0): mean 1 % of nc, 30 % rms."""
    return 1e25 * xp.clip(1 + 0.3 * f / f.std(), 0, None)


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.p, self.path = index, None, f"/tmp/tt_clocks_{os.getpid()}.csv"

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [x.strip() for x in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for n, v in zip(names, c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------- CPU arm
_CPU = {}


def _cpu_worker(args):
    """One Pool worker = one bundle, as example_multiprocess.py:41-51: seed, init_beam, solve (scipy RK45, default
    tolerances), shadowgraphy histogram.  kind "reference": the reference's OWN ElectronCube / Shadowgraphy objects
    (oracle/_ref, unmodified modules); kind "port": the numpy/scipy restatement (oracle/ref_numpy.py)."""
    i, n = args
    np.random.seed(1000 + i)
    if _CPU["kind"] == "reference":
        from oracle import reference_live as live
        pt, rtm, _ = live.load()
        cube = _CPU["cube"]
        cube.init_beam(n, BEAM_SIZE, DIVERGENCE)
        with live.NfevCounter(pt) as cnt:
            rf = live.quiet(cube.solve)
        sh = rtm.Shadowgraphy(rf)
        sh.solve()
        sh.histogram(bin_scale=10)
        return n, cnt.nfev * n, sh.H.sum()
    from oracle import ref_numpy as orc
    s0 = orc.init_beam(n, BEAM_SIZE, DIVERGENCE, EXTENT, "z")
    rf, sf, nfev = orc.solve(_CPU["field"], s0, EXTENT, "z")
    H, _, _ = orc.histogram(orc.detector("shadowgraphy", rf))
    return n, nfev, H.sum()


def host_grf_cube(n_half, seed=7):
    """Host-side synthetic ne cube of the same statistics (k^-11/3 GRF, M = 2N+1) for the CPU arm when
    no device cube is available: white noise shaped in Fourier space (scipy.fft, all host threads)."""
    import scipy.fft as sfft
    M = 2 * n_half + 1
    rng = np.random.default_rng(seed)
    w = rng.standard_normal((M, M, M), dtype=np.float32)
    F = sfft.rfftn(w, workers=-1)
    k = np.fft.fftfreq(M)
    kz = np.fft.rfftfreq(M)
    K2 = (k[:, None, None] ** 2 + k[None, :, None] ** 2 + kz[None, None, :] ** 2).astype(np.float32)
    K2[0, 0, 0] = 1.0
    F *= K2 ** (-11.0 / 12.0)          # sqrt(k^-11/3)
    F[0, 0, 0] = 0
    f = sfft.irfftn(F, s=(M, M, M), workers=-1)
    return ne_from_field(f.astype(np.float64), np)


def cpu_reference_setup(ne_host, M, prefer_reference=True):
    """The CPU arm's field: the reference's own ElectronCube (external_ne + calc_dndr, particle_tracker.py:212-241) when
    its modules are at hand (oracle/_ref on the GPU box, /root/reference in the build container), else the port.
    The C restatement borrows the gradient arrays of whichever was built (no copy)."""
    from oracle import reference_live as live
    x = np.linspace(-EXTENT, EXTENT, M)
    if prefer_reference and live.available():
        pt, _, where = live.load()
        cube = pt.ElectronCube(x, x, x, "z")
        cube.external_ne(ne_host)
        cube.calc_dndr(LWL)
        cube.ne = cube.ne_nc = None              # only the interpolators are used from here on
        _CPU.update(kind="reference", cube=cube, where=where, grid=(x, x, x), grads=(cube.dndx, cube.dndy, cube.dndz))
    else:
        from oracle import ref_numpy as orc
        f = orc.make_field(ne_host, x, x, x, LWL)
        _CPU.update(kind="port", field=f, where="oracle/ref_numpy.py", grid=f.ix.grid, grads=(f.ix.values, f.iy.values, f.iz.values))


def cpu_reference_step(rays_per_worker, cores, min_seconds=10.0, max_rounds=6):
    """One bounded sample over `cores` processes: rounds of rays_per_worker rays per process (new seeds
    each round) until at least min_seconds of wall time (the cost of scipy's global adaptive step depends
    strongly on the hardest ray of a bundle, so a fixed ray count gives anything from 1 s to minutes).
    Returns dict(rays, ray_rhs_evals, seconds, rounds)."""
    import multiprocessing as mp
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"                      # pinned: oversubscription costs 3x
    ctx = mp.get_context("fork")                 # cube built once before the fork (copy-on-write)
    tot = {"rays": 0, "ray_rhs_evals": 0, "seconds": 0.0, "rounds": 0}
    with ctx.Pool(cores) as pool:
        while tot["rounds"] < max_rounds and tot["seconds"] < min_seconds:
            t0 = time.perf_counter()
            out = pool.map(_cpu_worker, [(tot["rounds"] * cores + i, rays_per_worker) for i in range(cores)])
            tot["seconds"] += time.perf_counter() - t0
            tot["rays"] += sum(o[0] for o in out)
            tot["ray_rhs_evals"] += sum(o[1] for o in out)
            tot["rounds"] += 1
    return tot


def cpu_c_port_sample(cores, rays_per_bundle, min_seconds=3.0, max_rounds=40):
    """The same workload through the plain-C restatement of the reference path (oracle/tt_oracle.c: numpy.gradient,
    RegularGridInterpolator and solve_ivp RK45 written out, scipy's default tolerances, one adaptive step sequence
    per bundle like ElectronCube.solve), bundles handed to `cores` threads.  Reported next to the scipy figure as
    the stronger CPU baseline; the gradient arrays are shared with the scipy field (no copy)."""
    from oracle import c_oracle as orc_c
    from oracle import ref_numpy as orc
    x, y, z = _CPU["grid"]
    cf = orc_c.GradientField(x, y, z, *_CPU["grads"])
    tot = {"rays": 0, "ray_rhs_evals": 0, "seconds": 0.0, "rounds": 0}
    while tot["rounds"] < max_rounds and tot["seconds"] < min_seconds:
        n = cores * 4 * rays_per_bundle
        np.random.seed(5000 + tot["rounds"])
        s0 = orc.init_beam(n, BEAM_SIZE, DIVERGENCE, EXTENT, "z")
        t0 = time.perf_counter()
        rf, _, evals = orc_c.solve(cf, s0, EXTENT, "z", batch=rays_per_bundle, threads=cores)
        H, _, _ = orc.histogram(orc.detector("shadowgraphy", rf))
        tot["seconds"] += time.perf_counter() - t0
        tot["rays"] += n
        tot["ray_rhs_evals"] += evals
        tot["rounds"] += 1
    return {"value": tot["ray_rhs_evals"] / 4.0 / tot["seconds"], "unit": "ray-steps/s", "cores": cores, "kind": "port",
            "rays_per_s": tot["rays"] / tot["seconds"],
            "sample": f"C restatement (oracle/tt_oracle.c), {cores} threads x {4 * tot['rounds']} bundles of {rays_per_bundle} rays, "
                      f"RK45 default rtol=1e-3, ray-steps = nfev*rays/4; {tot['seconds']:.1f} s"}


def cpu_line_fields(res, cores, rays_per_worker, M, same_cube):
    steps = res["ray_rhs_evals"] / 4.0           # 4 RHS evaluations = 1 RK4-equivalent ray-step
    kind = _CPU["kind"]
    what = ("the reference's own ElectronCube.solve + Shadowgraphy (" + _CPU["where"] + ") under multiprocessing.Pool.map as "
            "example_multiprocess.py:41-51" if kind == "reference" else "numpy/scipy restatement (oracle/ref_numpy.py) under Pool.map")
    return {"value": steps / res["seconds"], "unit": "ray-steps/s", "cores": cores, "kind": kind,
            "rays_per_s": res["rays"] / res["seconds"], "rays_per_bundle": rays_per_worker,
            "cube": ("the GPU arm's own cube (same realisation)" if same_cube else
                     "host-generated k^-11/3 GRF cube of the same size and statistics (another realisation than the GPU arm's device-Philox cube)"),
            "sample": f"{what}; {cores} processes x {res.get('rounds', 1)} bundles of {rays_per_worker} rays on a {M}^3 "
                      f"cube; scipy RK45 default rtol=1e-3, ray-steps = nfev*rays/4; {res['seconds']:.1f} s"}


def run_reference_arm(args, wl):
    """--impl reference: the reference's CPU implementation alone, rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    n_half, _, desc = WORKLOADS[wl]
    M = 2 * n_half + 1
    cores = os.cpu_count() or 1
    rays_per_worker = args.cpu_rays
    if args.cube_file:                     # the GPU arm's own cube (cpu_baseline leg)
        ne = np.load(args.cube_file).astype(np.float64)
        M = ne.shape[0]
    else:
        ne = host_grf_cube(n_half)
    cpu_reference_setup(ne, M, prefer_reference=not args.cpu_port)
    del ne
    for _ in range(min(args.warmup, 1)):
        cpu_reference_step(max(rays_per_worker // 4, 50), cores, min_seconds=0.0, max_rounds=1)
    tot = {"rays": 0, "ray_rhs_evals": 0, "seconds": 0.0, "rounds": 0}
    for _ in range(args.steps):
        r = cpu_reference_step(rays_per_worker, cores, min_seconds=10.0 if args.steps == 1 else 5.0)
        for k in tot:
            tot[k] += r[k]
    cb = cpu_line_fields(tot, cores, rays_per_worker, M, bool(args.cube_file))
    try:
        cb_c = cpu_c_port_sample(cores, rays_per_worker)
    except Exception as e:                  # reported, never silently dropped
        cb_c = {"value": None, "unit": "ray-steps/s", "cores": cores, "kind": "port", "sample": f"failed: {type(e).__name__}: {e}"}
    line = {"metric": "ray-steps/s", "value": cb["value"], "unit": "ray-steps/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot["seconds"] / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "cores": cores,
            "rays_per_s": cb["rays_per_s"],
            "config": {"workload": desc, "cube": f"{M}^3", "rays_per_step": tot["rays"] // max(args.steps, 1),
                       "rays_per_bundle": rays_per_worker, "cube_realisation": cb["cube"]},
            "cpu_baseline": cb, "cpu_baseline_c": cb_c,
            "e2e": {"value": cb["value"], "unit": "ray-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------- GPU arm
TRACE_KERNELS = {       # (mangled-name substring, display name) of the kernel that dominates a step, per (dtype, aux, faces)
    ("float32", False, True): ("trace_face_kernel_f32x2ILb0", "trace_face_kernel_f32x2<false>"),
    ("float32", False, False): ("trace_event_kernel_f32x2ILb1ELb0ELb1ELb0", "trace_event_kernel_f32x2<1,0,1,0>"),
    ("float32", True, False): ("trace_event_kernel_f32x2ILb1ELb1ELb1ELb1", "trace_event_kernel_f32x2<1,1,1,1>"),
    ("float32", True, True): ("trace_face_aux_kernel_f32x2ILb0", "trace_face_aux_kernel_f32x2<false>"),
    ("float64", False, False): ("trace_event_kernelIdLb1", "trace_event_kernel<double,true>"),
}


def kernel_sass_sha(mangled_substr):
    """sha256 of the SASS text of one kernel of the LOADED libtt_b200.so (cuobjdump), or None.  profiles/traffic.json
    stamps every ncu-derived entry with it, so a stale instruction mix cannot be applied to a rebuilt kernel."""
    import hashlib
    import re
    import shutil
    from turbulence_tracing_b200 import _lib
    exe = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    try:
        out = subprocess.run([exe, "-sass", _lib.lib_path()], capture_output=True, text=True, timeout=120).stdout
    except Exception:
        return None
    h, on, found = hashlib.sha256(), False, False
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            on = mangled_substr in m.group(1)
            found = found or on
            continue
        if on:
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
            if m:
                h.update(m.group(1).strip().encode())
    return h.hexdigest() if found else None


def profile_entry(key):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)["entries"].get(key)
    except Exception:
        return None


def build_roofline(key, kern, steps_per_launch, kernel_ms, clocks, bytes_per_step, bytes_note):
    """Primary bound of the trace kernels: the FP32 (FMA) pipe.  The kernels keep the cell's polynomial in registers and
    load 48 B (one cell face) per ray-step from L1/L2, so HBM carries a few % of its bandwidth and an HBM fraction says
    nothing (it is reported as `hbm_algorithmic` for SURVEY section 8d, with the measured DRAM traffic beside it).
    achieved = ray-steps/s / 32 x (FMA-pipe issue slots per warp and ray-step, from the ncu opcode mix of THIS build:
    the entry in profiles/traffic.json carries the sha256 of the kernel's SASS, checked against the loaded library) x
    64 flop per slot (32 lanes x FMA); peak = 148 SMs x 4 sub-partitions x 1 slot per clock x sampled SM clock."""
    mangled, name = kern
    e = profile_entry(key) or {}
    sha = kernel_sass_sha(mangled)
    stale = None if sha is None or not e.get("sass_sha256") else (sha != e["sass_sha256"])
    peak_hbm, peak_src = measured_hbm_peak()
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    slots = e.get("fma_pipe_slots_per_warp_step")
    warp_steps_per_s = steps_per_launch / 32.0 / (kernel_ms * 1e-3)
    peak_slots = 148 * 4 * sm_mhz * 1e6
    r = {"bound": "fp32_pipe", "kernel": name, "unit": "TFLOP/s", "kernel_ms": kernel_ms,
         "ray_steps_per_launch": steps_per_launch,
         "peak": peak_slots * 64 / 1e12, "peak_source": f"148 SMs x 128 FP32 lanes x 2 flop x {sm_mhz:.0f} MHz (SM clock sampled during the timed region)",
         "achieved": None, "frac": None, "fma_pipe_slots_per_warp_step": slots, "profile": e.get("source"),
         "profile_matches_loaded_kernel": (None if stale is None else (not stale)), "kernel_sass_sha256": sha,
         "traffic": (e.get("dram_bytes_per_launch") / 1e9) if e.get("dram_bytes_per_launch") else None,
         "traffic_unit": "GB per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
         "ncu": {k: e.get(k) for k in ("ncu_kernel_ms", "fma_pipe_busy_pct", "issue_slots_busy_pct", "l1_hit_pct", "l2_hit_pct",
                                       "dram_pct_of_peak", "warp_instructions_per_warp_step", "registers", "warps_active_pct",
                                       "eligible_warps_per_scheduler") if k in e}}
    if slots:
        ach = warp_steps_per_s * slots
        r["achieved"] = ach * 64 / 1e12
        r["frac"] = ach / peak_slots
        r["note"] = ("achieved = FMA-pipe issue slots/s x 64 flop (a packed FP32x2 instruction holds the pipe for 2 slots); "
                     "frac = share of the FP32 pipe's issue cycles in use = ncu sm__pipe_fma_cycles_active")
        om = e.get("operand_model")
        if om and not stale:
            # cycles per warp-step and sub-partition, live: kernel time x SM clock x (148 SMs x 4) / warp-steps
            cyc = kernel_ms * 1e-3 * sm_mhz * 1e6 * 148 * 4 / (steps_per_launch / 32.0)
            r["operand_bandwidth_model"] = {"cycles_per_warp_step_model": om["cycles_per_warp_step"], "cycles_per_warp_step_measured": cyc,
                                            "frac": om["cycles_per_warp_step"] / cyc, "mix": om["mix"], "note": om["note"]}
    alg = steps_per_launch * bytes_per_step / (kernel_ms * 1e-3) / 1e9
    r["hbm_algorithmic"] = {"bound": "hbm", "achieved": alg, "peak": peak_hbm, "unit": "GB/s", "frac": alg / peak_hbm,
                            "algorithmic_bytes_per_ray_step": bytes_per_step, "peak_source": peak_src,
                            "note": bytes_note + "; rays that share a cell are served by L1/L2, DRAM carries `traffic`, so this "
                                                 "fraction exceeds 1 and is no bound of the kernel"}
    return r


class GpuCtx:
    pass


def gpu_measure(ctx, args, wl, *, rays_per_rank, first_ray, dtype, steps, warmup, do_e2e, sampler=None, beam=BEAM_SIZE,
                bin_scale=10, steps_per_cell=1, face_grid=True, pageable=False, ne_cache=None):
    """One measurement: `steps` timed passes of the hot path over this rank's `rays_per_rank` rays of workload `wl`.
    Returns a dict (same on every rank for the reduced figures)."""
    torch, ttd, pt, rtm, tg, _lib = ctx.torch, ctx.ttd, ctx.pt, ctx.rtm, ctx.tg, ctx.lib
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    n_half, _, desc = WORKLOADS[wl]
    M = 2 * n_half + 1
    x = np.linspace(-EXTENT, EXTENT, M)
    lib = _lib.load()

    # ---- synthetic inputs: GRF cube on rank 0 (device Philox + cuFFT), NCCL broadcast ------------
    t0 = time.perf_counter()
    if ne_cache is not None and ne_cache.get("M") == M:
        ne = ne_cache["ne"]
    else:
        if ne_cache is not None:
            ne_cache.clear()
            torch.cuda.empty_cache()
        ne = torch.empty((M, M, M), dtype=torch.float32, device=dev)
        if rank == 0:
            f = tg.gaussian3D_FFT(n_half, SPECTRUM, seed=1234, dtype="float32", return_device=True).torch
            ne.copy_(ne_from_field(f, torch))
            del f
        ttd.broadcast_cube(ne, src=0)
        if ne_cache is not None:
            ne_cache.update(M=M, ne=ne)
    torch.cuda.synchronize()
    t_cube = time.perf_counter() - t0

    aux = wl == "c4"
    if aux:                                    # magnetised / absorbing plasma (parity unpinned)
        Bvec = torch.zeros((M, M, M, 3), dtype=torch.float32, device=dev)
        Te = torch.empty((M, M, M), dtype=torch.float32, device=dev)
        if rank == 0:
            g1 = tg.gaussian3D_FFT(n_half, SPECTRUM, seed=77, dtype="float32", return_device=True).torch
            Bvec[..., 2] = 10.0
            Bvec[..., 0] = 2.0 * g1 / g1.std()
            Te.copy_(100.0 * torch.clamp(1 + 0.2 * g1 / g1.std(), min=0.1))
            del g1
        ttd.broadcast_cube(Bvec, src=0)
        ttd.broadcast_cube(Te, src=0)

    def make_cube():
        c = pt.ElectronCube(x, x, x, dtype=dtype, steps_per_cell=steps_per_cell, keep_sf=False, verbose=False,
                            B_on=aux, inv_brems=aux, phaseshift=aux, face_grid="auto" if face_grid else False)
        c.kernel_variant = args.variant
        if aux:
            c.external_B(Bvec)
            c.external_Te(Te)
            c.external_Z(1.0)
        return c

    rays = rays_per_rank
    cube = make_cube()
    cube.external_ne(ne)
    cube.init_beam(rays, beam, DIVERGENCE, seed=99, first_ray=first_ray)
    s0_dev = cube.s0
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    trace_ms = []
    hist_ms = []               # detector kernel (optics + histogram), one launch per step
    phase_events = []          # per step: events at the phase boundaries (read after the timed region)

    def mark(lst):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        lst.append(e)

    def detector(rf, to_host):
        sh = rtm.Shadowgraphy(rf)
        sh.solve()
        sh.histogram(bin_scale=bin_scale, to_host=to_host)
        return sh

    def step_device():
        marks = []
        mark(marks)
        cube.external_ne(ne)
        cube.calc_dndr(LWL)
        mark(marks)
        cube.s0 = s0_dev
        rf = cube.solve()
        mark(marks)
        sh = detector(rf, False)
        mark(marks)
        H = ttd.allreduce_histograms([sh.H_dev])[0]
        mark(marks)
        phase_events.append(marks)
        return H, cube._steps_dev

    def timed(fn, c, warmup, steps, sampler=None):
        for _ in range(warmup):
            fn()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        c._trace_events = []                    # CUDA events around every trace launch from here on
        rtm._kernel_events = []                 # ... and around every detector-kernel launch
        if sampler:
            sampler.start()
        l0 = int(lib.tt_launch_count())
        ev[0].record()
        counters, last = [], None
        for _ in range(steps):
            last = fn()
            counters.append(last[1])
        ev[1].record()
        torch.cuda.synchronize()
        launches = int(lib.tt_launch_count()) - l0
        clocks = sampler.stop() if sampler else None
        if world > 1:
            torch.distributed.barrier()
        ms = ev[0].elapsed_time(ev[1])
        trace_ms[:] = c.trace_ms()
        c._trace_events = None
        hist_ms[:] = [a.elapsed_time(b) for a, b in rtm._kernel_events]
        rtm._kernel_events = None
        ms = ttd.allreduce_scalar(float(ms), "max", device=dev)
        tot_steps = ttd.allreduce_scalar(int(sum(int(t.item()) for t in counters)), "sum", device=dev)
        launches = ttd.allreduce_scalar(int(launches), "sum", device=dev)
        return ms, tot_steps, last, clocks, launches

    ms, tot_steps, last, clocks, launches = timed(step_device, cube, warmup, steps, sampler)
    kernel_ms = float(np.mean(trace_ms)) if trace_ms else float("nan")
    names = ["calc_dndr", "face_grid+sort+trace", "optics+hist", "allreduce"]
    phases = {n: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in phase_events[-steps:]])) for i, n in enumerate(names)}
    phases["trace_kernel"] = kernel_ms
    phases["detector_kernel"] = float(np.mean(hist_ms)) if hist_ms else float("nan")
    if world > 1:                        # every rank's breakdown (rank skew shows up as all-reduce wait)
        allp = [None] * world
        torch.distributed.all_gather_object(allp, phases)
        phases = allp
    rays_total = ttd.allreduce_scalar(int(rays), "sum", device=dev)
    H_dev = last[0]
    # ---- validity of the timed work (last step): every ray marched to the far face by the event kernel, none lost ----
    st = cube.status.torch
    n_exit = ttd.allreduce_scalar(int((st == 1).sum().item()), "sum", device=dev)
    # a terminal state: far face / side face / time cap / never entered (the GENERAL bit only says which kernel finished it;
    # 255 = handed to the second pass and never finished)
    n_term = ttd.allreduce_scalar(int((((st & 15) != 0) & (st != 255)).sum().item()), "sum", device=dev)
    hist_sum = int(H_dev.sum().item())
    expect = rays_total * (M - 1) * steps_per_cell * steps
    fitted = beam <= BEAM_SIZE           # the beam fits the cube's cross-section: every ray must march to the far face
    checks = {"rays": rays_total, "status_exit_face": n_exit, "status_terminal": n_term, "histogram_sum": hist_sum,
              "ray_steps": tot_steps, "ray_steps_expected_if_all_marched": expect,
              "rule": ("every ray at the far face, ray-steps = rays x planes, 0 < histogram <= rays" if fitted else
                       "wide beam: every ray in a terminal state (far face, side face or never entered), "
                       "0.95 x rays x planes <= ray-steps <= rays x planes, 0 < histogram <= rays"),
              "ok": bool((n_exit == rays_total and tot_steps == expect if fitted else
                          n_term == rays_total and 0.95 * expect <= tot_steps <= expect) and 0 < hist_sum <= rays_total)}
    used_faces = cube._faces is not None
    out = {"value": tot_steps / (ms * 1e-3), "ms_per_step": ms / steps, "rays_per_s": rays_total * steps / (ms * 1e-3),
           "rays_total": rays_total, "rays_per_rank": rays, "phases_ms": phases, "kernel_ms": kernel_ms, "clocks": clocks,
           "gpu_launches": launches, "checks": checks, "cube_setup_s": t_cube, "M": M, "aux": aux, "faces": used_faces,
           "steps_per_launch": tot_steps / (steps * world), "desc": desc, "e2e": None, "ne": ne}

    # ---- e2e: public API, host buffers in, histogram + counter out ----------------------------------
    if do_e2e:
        def host_buffers(pinned):
            ne_h = torch.empty((M, M, M), dtype=torch.float32, pin_memory=pinned)
            ne_h.copy_(ne)
            s0_h = torch.empty((6, rays), dtype=torch.float64, pin_memory=pinned)
            s0_h.copy_(s0_dev.torch)
            torch.cuda.synchronize()
            return ne_h, s0_h

        def run_e2e(pinned, steps_e, warm_e):
            ne_host, s0_host = host_buffers(pinned)
            cube2 = make_cube()
            if args.e2e_chunk:
                cube2.pipeline_chunk_rays = args.e2e_chunk
            ne_np, s0_np = ne_host.numpy(), s0_host.numpy()
            e2e_events = []

            def step_e2e():
                marks = []
                mark(marks)
                if world > 1:                           # each rank uploads 1/world of the cube, one all-gather over NVLink
                    cube2.external_ne(ttd.upload_cube_sharded(ne_np, device=dev))
                else:
                    cube2.external_ne(ne_np)            # host numpy -> H2D inside calc_dndr
                cube2.calc_dndr(LWL)
                mark(marks)
                cube2.s0 = s0_np                        # host numpy -> H2D inside solve
                rf = cube2.solve()
                mark(marks)
                sh = detector(rf, True)                 # D2H of the histogram inside
                H = ttd.allreduce_histograms([sh.H_dev])[0]
                _ = cube2.ray_steps                     # D2H of the counter
                mark(marks)
                e2e_events.append(marks)
                return H, cube2._steps_dev

            ms2, tot2, last2, _, launches2 = timed(step_e2e, cube2, warm_e, steps_e)
            e2e_names = ["h2d_cube+calc_dndr", "h2d_rays+face_grid+sort+trace", "optics+hist+d2h"]
            ph = {n: float(np.mean([m[i].elapsed_time(m[i + 1]) for m in e2e_events[-steps_e:]])) for i, n in enumerate(e2e_names)}
            if world > 1:
                allp = [None] * world
                torch.distributed.all_gather_object(allp, ph)
                ph = allp
            r = {"value": tot2 / (ms2 * 1e-3), "unit": "ray-steps/s", "ms_per_step": ms2 / steps_e, "steps": steps_e, "phases_ms": ph,
                 "host_memory": "pinned (torch pin_memory)" if pinned else "pageable (plain numpy arrays, as a drop-in caller passes)",
                 "trace_launches_per_step": len(trace_ms) // max(steps_e, 1),
                 "trace_kernels_ms_per_step": float(np.sum(trace_ms)) / max(steps_e, 1),
                 "h2d_bytes_per_step": int(-(-ne_host.numel() // world) * 4 + s0_host.numel() * 8),      # per rank
                 "d2h_bytes_per_step": int(last2[0].numel() * 8 + 8), "gpu_launches": launches2,
                 "pipeline": {"chunk_cap_rays": getattr(cube2, "pipeline_chunk_rays", None) or "adaptive",
                              "first_chunk_upload_gbs": getattr(cube2, "last_upload_gbs", None),
                              "best_chunk_upload_gbs": getattr(cube2, "last_upload_gbs_max", None),
                              "chunk_growth": getattr(cube2, "last_pipeline_growth", None)},
                 "cube_upload": "sharded: 1/N per rank over PCIe + one NCCL all-gather (distributed.upload_cube_sharded)" if world > 1 else "whole cube",
                 "api": "ElectronCube.external_ne/calc_dndr/solve + Shadowgraphy.solve/histogram, numpy in, H out"}
            del ne_host, s0_host, cube2
            return r

        out["e2e"] = run_e2e(True, steps, max(2, min(warmup, 3)))
        if pageable:
            out["e2e_pageable"] = run_e2e(False, max(1, steps // 2), 1)
    del cube
    return out


def compact(m):
    """sub-record of the JSON line from a gpu_measure() result"""
    r = {k: m[k] for k in ("value", "ms_per_step", "rays_per_s", "rays_total", "kernel_ms", "phases_ms", "gpu_launches", "checks")}
    r["unit"] = "ray-steps/s"
    if m.get("e2e"):
        r["e2e"] = {k: m["e2e"][k] for k in ("value", "ms_per_step", "phases_ms", "h2d_bytes_per_step", "d2h_bytes_per_step", "host_memory")}
    return r


def run_gpu_arm(args, wl):
    import torch
    from turbulence_tracing_b200 import distributed as ttd
    from turbulence_tracing_b200 import particle_tracker as pt, ray_transfer_matrix as rtm, turboGen as tg
    from turbulence_tracing_b200 import _lib

    ctx = GpuCtx()
    ctx.torch, ctx.ttd, ctx.pt, ctx.rtm, ctx.tg, ctx.lib = torch, ttd, pt, rtm, tg, _lib
    ctx.rank, ctx.world, local = ttd.init_from_env()
    rank, world = ctx.rank, ctx.world
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    ctx.dev = dev = torch.device("cuda", local)
    numa = ttd.bind_to_gpu_numa_node(local) if hasattr(ttd, "bind_to_gpu_numa_node") else None
    _lib.load(build_if_missing=False)
    n_half, rays, desc = WORKLOADS[wl]
    if args.rays:
        rays = args.rays
    dtype = args.dtype
    cache = {}

    # ---- the headline record: weak scaling, `rays` per GPU -----------------------------------------------------
    first, _ = ttd.shard_range(rays * world, rank, world)
    main = gpu_measure(ctx, args, wl, rays_per_rank=rays, first_ray=first, dtype=dtype, steps=args.steps, warmup=args.warmup,
                       do_e2e=not args.no_e2e, sampler=ClockSampler(local), steps_per_cell=args.steps_per_cell,
                       face_grid=not args.no_face_grid, pageable=(not args.no_e2e and not args.no_extras), ne_cache=cache,
                       beam=args.beam or BEAM_SIZE)
    M, aux, clocks = main["M"], main["aux"], main["clocks"]
    faces = main["faces"]
    key = f"{wl}/{dtype}/" + ("faces" if faces else str(args.variant or 3))
    kern = TRACE_KERNELS.get((dtype, aux, faces), ("", "trace kernel"))
    if faces and aux:
        bps, note = 128, "one cell face of both coefficient grids per ray-step: 3 + 5 words of 16 B (gradient; ne/nc, B, kappa)"
    elif faces:
        bps, note = 48, "one cell face = 3 x 16 B of bilinear coefficients per ray-step (SURVEY 8d counted 4 stages x 8 corners x 16 B = 512 B for a per-stage gather)"
    else:
        bps = 64 * (2 if dtype == "float64" else 1) * (2 if aux else 1)
        note = "4 corners of the next plane per ray-step, 16 B each in FP32 (32 B in FP64; x2 with the B/kappa grid)"
    roofline = build_roofline(key, kern, main["steps_per_launch"], main["kernel_ms"], clocks, bps, note) if not args.rays else None
    if roofline is not None:
        # the second kernel of the step, HBM-bound by design: 32 B per ray read once (x, theta, y, phi in FP64), image in shared memory
        ph = main["phases_ms"][0] if isinstance(main["phases_ms"], list) else main["phases_ms"]
        dms = ph.get("detector_kernel")
        if dms and dms == dms:
            peak_hbm, peak_src = measured_hbm_peak()
            gbs = 32.0 * main["rays_per_rank"] / (dms * 1e-3) / 1e9
            roofline["detector_kernel"] = {"bound": "hbm", "kernel": "optics_hist_smem16_kernel (optics_hist_kernel when the image does not fit shared memory)",
                                           "kernel_ms": dms, "algorithmic_bytes_per_ray": 32, "achieved": gbs, "peak": peak_hbm,
                                           "unit": "GB/s", "frac": gbs / peak_hbm, "peak_source": peak_src,
                                           "note": "CUDA events around the launch inside the timed region (rank 0)"}
            try:
                te = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["entries"].get(f"detector/{wl}/smem16")
            except Exception:
                te = None
            if te:                                   # (the main record bins at the default bin_scale = 10)
                roofline["detector_kernel"].update(traffic=te["dram_gb_per_launch"] * main["rays_per_rank"] / te["rays_per_launch"],
                                                   traffic_unit="GB per launch (dram__bytes_read.sum + dram__bytes_write.sum, ncu --set full)",
                                                   profile=te["profile"])

    # ---- sub-records (same machinery, fewer steps): strong scaling, configs[4] share, configs[1], configs[3], FP64, ... ----
    extra = {}
    if not args.no_extras and wl == "c3" and not args.rays:
        xs, xw = max(2, args.steps // 2), 2
        if world > 1:
            f0, cnt = ttd.shard_range(WORKLOADS["c3"][1], rank, world)
            m = gpu_measure(ctx, args, "c3", rays_per_rank=cnt, first_ray=f0, dtype="float32", steps=xs, warmup=xw,
                            do_e2e=not args.no_e2e, ne_cache=cache)
            extra["strong_c3"] = dict(compact(m), scaling="strong",
                                      workload=f"configs[2] strong scaling: 513^3 cube, 1e8 rays TOTAL sharded over {world} GPUs "
                                               f"(example_MPI.py:117-149), histogram all-reduce")
        if world == 1:
            m = gpu_measure(ctx, args, "c3", rays_per_rank=rays, first_ray=0, dtype="float32", steps=xs, warmup=xw, do_e2e=False,
                            ne_cache=cache, beam=EXTENT)
            extra["wide_beam_c3"] = dict(compact(m), workload="configs[2] with beam_size = extent (example_multiprocess.py:46-50): "
                                         "rays outside the cube's cross-section or leaving sideways take the general kernel",
                                         deferred_fraction=1.0 - m["checks"]["status_exit_face"] / m["checks"]["rays"])
            m = gpu_measure(ctx, args, "c3", rays_per_rank=rays, first_ray=0, dtype="float32", steps=xs, warmup=xw, do_e2e=False,
                            ne_cache=cache, bin_scale=1)
            extra["bin_scale_1_c3"] = dict(compact(m), workload="configs[2] with the detector at bin_scale=1 (2574 x 3448 bins, example_MPI.py:69)")
            m = gpu_measure(ctx, args, "c3", rays_per_rank=rays, first_ray=0, dtype="float32", steps=xs, warmup=xw, do_e2e=False,
                            ne_cache=cache, face_grid=False)
            extra["corner_grid_c3"] = dict(compact(m), workload="configs[2] through tt_trace (float4 corner grid, round-1 path)")
            for name, w, dt, spc in (("c2", "c2", "float32", 1), ("c2_fp64", "c2", "float64", 2), ("c4", "c4", "float32", 1)):
                m = gpu_measure(ctx, args, w, rays_per_rank=WORKLOADS[w][1], first_ray=0, dtype=dt, steps=xs, warmup=xw,
                                do_e2e=False, ne_cache=cache, steps_per_cell=spc)
                extra[name] = dict(compact(m), workload=WORKLOADS[w][2] + f", {dt}, {spc} step(s) per cell")
        if world in (1, 8):
            per = WORKLOADS["c5"][1]
            f0, _ = ttd.shard_range(per * world, rank, world)
            main.pop("ne", None)
            # 1025^3: 17 GB node grid + 52 GB face grid + 10 GB of rays.  Hand the cached blocks of the smaller workloads
            # back first: carved out of a fragmented pool the same kernel ran 980 instead of 793 ms (measured; alone, as
            # `--workload c5`, and at 8 GPUs behind only one other sub-record: 791-793 ms)
            cache.clear()
            import gc
            gc.collect()
            torch.cuda.empty_cache()
            m = gpu_measure(ctx, args, "c5", rays_per_rank=per, first_ray=f0, dtype="float32", steps=2, warmup=1,
                            do_e2e=(not args.no_e2e and world == 8), ne_cache=cache)
            extra["c5"] = dict(compact(m), scaling="weak", workload=WORKLOADS["c5"][2] +
                               (" -- all 8 shares: 1e9 rays" if world == 8 else " -- one GPU's share of the 8-GPU job"))
    cache.clear()
    main.pop("ne", None)

    # ---- CPU baseline on a bounded sample of the same cube (rank 0, N = 1 only) ----------------------
    cpu = cpu_c = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            # a fresh process (no CUDA context to fork) runs the CPU path on the very same cube
            n_half = WORKLOADS[wl][0]
            f = tg.gaussian3D_FFT(n_half, SPECTRUM, seed=1234, dtype="float32", return_device=True).torch
            path = f"/tmp/tt_ne_{os.getpid()}.npy"
            np.save(path, ne_from_field(f, torch).cpu().numpy())
            del f
            out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--workload", wl,
                                  "--cube-file", path, "--steps", "1", "--warmup", "0", "--cpu-rays",
                                  str(args.cpu_rays)], capture_output=True, text=True, timeout=1500)
            os.remove(path)
            ref_line = json.loads(out.stdout.strip().splitlines()[-1])
            cpu, cpu_c = ref_line["cpu_baseline"], ref_line.get("cpu_baseline_c")
        except Exception as e:      # reported, never silently dropped
            cpu = {"value": None, "unit": "ray-steps/s", "cores": os.cpu_count(), "kind": "reference",
                   "sample": f"failed: {type(e).__name__}: {e}"}

    if rank == 0:
        e2e = main["e2e"]
        line = {
            "metric": "ray-steps/s", "value": main["value"], "unit": "ray-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32" if dtype == "float32" else "f64", "data": "synthetic",
            "rays_per_s": main["rays_per_s"], "host_cores": os.cpu_count(),
            "config": {"workload": desc, "cube": f"{M}^3", "rays_per_gpu": rays, "steps_per_cell": args.steps_per_cell,
                       "beam_size_m": args.beam or BEAM_SIZE, "divergence_rad": DIVERGENCE, "detector": "Shadowgraphy bin_scale=10",
                       "parallelism": f"rays sharded x{world}, cube replicated (NCCL broadcast), histogram all-reduce",
                       "trace_path": "face-coefficient grid (tt_build_face_grid + tt_trace_faces)" if faces else "float4 corner grid (tt_trace)",
                       "l2": "inputs larger than L2 (gradient grid %.2f GB%s, rays %.2f GB)" % (
                           M**3 * (16 if dtype == "float32" else 32) / 1e9,
                           (" + face-coefficient grid %.2f GB" % ((M - 1) ** 2 * (M + 1) * 48 / 1e9)) if faces else "", rays * 48 / 1e9),
                       "cube_setup_s": main["cube_setup_s"], "kernel_variant": args.variant, "numa_binding": numa},
            "e2e": e2e, "e2e_pageable": main.get("e2e_pageable"),
            "gpu_launches": main["gpu_launches"],
            "gpu_launches_note": "counted by the library (tt_launch_count: every <<<>>> of its own kernels) over the timed region, all ranks; "
                                 "per step and GPU: calc_dndr_kernel, face_grid_kernel, morton_key_kernel, trace_face_kernel_f32x2, "
                                 "trace_kernel (second pass over deferred rays, returns at once when there are none), optics_hist_kernel "
                                 "(CUB's radix-sort passes are library kernels and not counted)",
            "roofline": roofline, "cpu_baseline": cpu, "cpu_baseline_c": cpu_c, "clocks": clocks, "phases_ms": main["phases_ms"],
            "checks": main["checks"], "histogram_sum": main["checks"]["histogram_sum"], "extra": extra,
        }
        emit(line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Everything that native libraries print to fd 1 (e.g. "NCCL version ...") goes to stderr; the ONE JSON
    line of the contract is written to the real stdout by emit()."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--rays", type=int, default=0, help="override rays per GPU")
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--steps-per-cell", type=int, default=1)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--cpu-rays", type=int, default=1000, help="rays per host process in the CPU baseline sample")
    ap.add_argument("--cube-file", default="", help="(reference arm) .npy ne cube to trace instead of a host GRF")
    ap.add_argument("--e2e-chunk", type=int, default=0, help="rays per upload chunk of the pipelined host-ray path (0 = library default)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-face-grid", action="store_true", help="trace over the float4 corner grid (tt_trace) instead of the face-coefficient grid")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sub-records (strong scaling, configs[1]/[3]/[4], FP64, wide beam, bin_scale=1, pageable e2e)")
    ap.add_argument("--beam", type=float, default=0.0, help="beam radius in m (default 4e-3)")
    ap.add_argument("--cpu-port", action="store_true", help="(reference arm) time the numpy/scipy restatement instead of the reference's own modules")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference_arm(args, args.workload)
    else:
        run_gpu_arm(args, args.workload)


if __name__ == "__main__":
    main()
