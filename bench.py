#!/usr/bin/env python
"""bench.py -- headline benchmark of the DepthFirst hot path (BASELINE.json metric, config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|C5v|...]

One "step" = one 1920x1080 mono frame of a 1M-Gaussian SH3 float16 synthetic cloud (SURVEY.md 8(d) recipe)
through gsm_render (C ABI). Prints ONE JSON line (rank 0):
  value        frames/s with inputs resident in HBM, per-step CUDA events on the launching stream,
               L2 flushed between steps; whole job = sum over ranks (views shard with no collective)
  e2e          the same frame through gsm_render_host_async/_wait: pinned host inputs -> H2D -> render -> D2H every step,
               two frames in flight on two renderers (blocking gsm_render_host reported beside it)
  roofline     dominant kernel: algorithmic bytes / its measured duration vs the
               measured HBM peak of MEASURED_PEAKS.json; stage_roofline lists every stage
  cpu_baseline the CPU oracle (a port of the reference's Metal kernels) timed on this box's cores (rank 0, N=1)
  modes        (bench_modes.py) the multi-GPU shard modes of SURVEY.md 8(e) measured in the same run: c3_strips (6 M, 3840x2160,
               one frame split by strips over the ranks, exchange = the routing kernel's NVLink peer stores, assembled image
               checked against rank 0's single-GPU frame), c5_views (256 orbit poses of 3 M at 720p split by view), c4_eyes
               (stereo, one eye per GPU). At N=1 they report the single-GPU numbers the N>1 efficiencies are relative to.
--impl reference times that CPU implementation alone on the same config (the Metal reference cannot run here).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# stdout carries exactly one JSON line: native libraries that print to fd 1 (NCCL's version banner under NCCL_DEBUG=VERSION)
# are sent to stderr for the whole run; emit_line() writes the result to the real stdout
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit_line(line: dict) -> None:
    sys.stdout.flush()
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from gsm_renderer_b200 import synthetic as syn  # noqa: E402

WORKLOADS = {
    # name: (N, sh_degree, precision, W, H, scale_median, description)
    "C1": (50_000, 1, "float32", 1920, 1080, 0.015, "C1: 50k Gaussians SH1 float32, 1920x1080 mono"),
    "C2": (1_000_000, 3, "float16", 1920, 1080, 0.015, "C2: 1M Gaussians SH3 float16, 1920x1080 mono"),
    "C5v": (3_000_000, 3, "float16", 1280, 720, 0.012, "C5 (one view): 3M Gaussians SH3 float16, 1280x720"),
    "C3": (6_000_000, 3, "float16", 3840, 2160, 0.008, "C3 (one GPU): 6M Gaussians SH3 float16, 3840x2160 mono"),
}
NEAR, FAR = 0.1, 100.0  # PLYBenchmarkTests.swift:60-62
KERNELS_PER_FRAME = 10  # project, compaction (+header), depth bucket rank + scatter + local sort, scan+expand, one tile onesweep pass + chunk count + chunk place (writes the ranges), blend


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def stage_bytes(N, V, Vp, I, T, P, rec_bytes, sh_bytes):
    """Algorithmic bytes per stage, BASELINE.md section 4 (compaction is fused into project: 0 extra)."""
    return {
        "project": N * (rec_bytes + 24) + Vp * sh_bytes + 16 * V,
        "depthSort": 68 * V,
        "expand": 60 * V + 6 * I,   # apply-order + prefix sum (20 V) run inside the expansion kernel: their bytes belong to it
        "tileSort": 26 * I,
        "ranges": 2 * I + 12 * T,
        "blend": 20 * I + 10 * P,
    }


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (NVML every 2 ms; nvidia-smi as fallback)."""

    def __init__(self, gpu_index=0):
        self.gpu, self.sm, self.mx, self.reasons, self.power = gpu_index, [], None, [], []
        self.ts = []
        self._stop = threading.Event()
        self.t = None
        self.src = "nvml"

    def _loop_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.gpu)
        self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        bits = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        while not self._stop.is_set():
            try:
                c = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(get_reasons(h))
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                self.ts.append(time.perf_counter()); self.sm.append(c); self.power.append(pw)
                self.reasons.append([n for b, n in bits.items() if r & b])
            except Exception:
                pass
            time.sleep(0.001)

    def _loop_smi(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.ts.append(time.perf_counter()); self.sm.append(float(out[0])); self.power.append(0.0)
                self.mx = float(out[1])
                self.reasons.append([n for n, v in zip(names, out[2:6]) if v.strip().lower().startswith("active")])
            except Exception:
                time.sleep(0.05)

    def start(self):
        try:
            import pynvml  # noqa: F401
            target = self._loop_nvml
        except Exception:
            target, self.src = self._loop_smi, "nvidia-smi"
        self.t = threading.Thread(target=target, daemon=True)
        self.t.start()

    def stop(self, t0=None, t1=None):
        self._stop.set()
        if self.t:
            self.t.join(timeout=3)
        idx = [i for i, t in enumerate(self.ts) if (t0 is None or t >= t0) and (t1 is None or t <= t1)]
        sm = [self.sm[i] for i in idx]
        reasons = sorted({n for i in idx for n in self.reasons[i]})
        pw = [self.power[i] for i in idx]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.mx, "reasons": reasons,
                "samples": len(sm), "samples_total": len(self.ts), "source": self.src,
                "power_w_max": max(pw) if pw else None}


def build_workload(name, seed=42):
    N, deg, prec, W, H, sm, desc = WORKLOADS[name]
    cloud = syn.synthetic_cloud(N, deg, seed=seed, scale_median=sm)
    g, h = cloud.pack(prec)
    return cloud, np.ascontiguousarray(g), np.ascontiguousarray(h), (N, deg, prec, W, H, desc)


def base_config(spec, V, I):
    """The `config` keys both arms print (the driver compares them)."""
    return {"workload": spec[5], "seed": 42, "near": NEAR, "far": FAR, "N": spec[0], "V": int(V), "I": int(I)}


def oracle_frame_runner(g, h, spec, threads=None):
    """Returns (run(), frame) for the CPU implementation of the whole frame (oracle/gsm_oracle.c)."""
    from oracle import binding as ob
    ob.build()
    N, deg, prec, W, H, _ = spec
    # all the host cores, whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1 to every rank)
    ob.lib().gsmo_set_num_threads(int(threads) if threads else (len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()))
    proj = syn.make_projection_matrix(W, H, NEAR, FAR)
    cam = ob.make_camera(np.eye(4), proj, (0, 0, 0), W, H, NEAR, FAR, syn.SH_COEFFS[deg], N, False)
    fr = ob.OracleFrame(N, W, H)
    p = ob.F16 if prec == "float16" else ob.F32

    def run():
        t = time.perf_counter()
        fr.render_mono(g, h, p, cam, W, H)
        return time.perf_counter() - t
    return run, fr, ob


def cpu_baseline_dict(fps, cores, sample, fr):
    st = {k: 1e3 * v for k, v in fr.stage_seconds.items()}
    total = st.get("total") or sum(v for k, v in st.items() if k != "total")
    share = st.get("clearBlend", 0.0) / total if total else None
    return {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample, "stage_ms": st,
            "blend_share": share,
            "note": "the port emulates binary16 arithmetic in software (one correctly rounded operation at a time): its clear+blend "
                    "stage is that emulation's cost, so this is a parity definition timed, not a tuned CPU renderer"}


def library_sort_bar(torch, V, I, T, stage_ms):
    """cub::DeviceRadixSort::SortPairs (tools/cub_bar.cu, built by __graft_entry__.build) on arrays of the frame's own sizes and
    key distributions, timed with CUDA events beside the frame's sort stages: V (depth key u32, gid) pairs over all 32 bits,
    I (tile id u16, instance) pairs over the id's bits. Measurement infrastructure only -- the product never loads it."""
    import ctypes as C
    lib = os.path.join(ROOT, "tools", "bin", "libcub_bar.so")
    if not os.path.exists(lib):
        return {"unavailable": "tools/bin/libcub_bar.so not built"}
    try:
        cub = C.CDLL(lib)
        cub.cub_sort_pairs.restype = C.c_int
        cub.cub_sort_pairs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                       C.c_int, C.POINTER(C.c_float)]
        rng = np.random.default_rng(0)
        out = {}
        depth = rng.uniform(2.0, 20.0, V).astype(np.float32)
        k32 = torch.from_numpy((depth.view(np.uint32) | np.uint32(0x80000000)).view(np.int32)).cuda()
        k16 = torch.from_numpy(rng.integers(0, T, I, dtype=np.int64).astype(np.int16)).cuda()
        tile_bits = max(1, int(T - 1).bit_length())
        for name, keys, bits, end_bit, ours in (("depthSort", k32, 32, 32, stage_ms.get("depthSort")),
                                                ("tileSort", k16, 16, tile_bits, stage_ms.get("tileSort", 0.0) + stage_ms.get("ranges", 0.0))):
            vals = torch.arange(keys.numel(), dtype=torch.int32, device="cuda")
            ko, vo, ms = torch.empty_like(keys), torch.empty_like(vals), C.c_float(0)
            rc = cub.cub_sort_pairs(keys.data_ptr(), vals.data_ptr(), ko.data_ptr(), vo.data_ptr(), keys.numel(), bits, 0, end_bit,
                                    torch.cuda.current_stream().cuda_stream, 20, C.byref(ms))
            if rc != 0:
                return {"unavailable": f"cub_sort_pairs rc {rc}"}
            out[name] = {"pairs": int(keys.numel()), "cub_us": ms.value * 1e3, "ours_us": None if ours is None else ours * 1e3,
                         "end_bit": end_bit}
        out["note"] = ("cub::DeviceRadixSort::SortPairs alone on resident arrays (L2 warm, 20 repetitions) vs this frame's stage times "
                       "(tileSort includes the tile ranges, which the MSD tile sort writes itself); cub sorts only the id's significant bits")
        return out
    except Exception as e:  # measurement extra: never fail the bench line
        return {"unavailable": repr(e)}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores (CPU port; Metal cannot run here)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cloud, g, h, spec = build_workload(args.workload)
    run, fr, ob = oracle_frame_runner(g, h, spec)
    cores = ob.lib().gsmo_num_threads()
    for _ in range(max(1, min(args.warmup, 2))):
        run()
    times = [run() for _ in range(args.steps)]
    ms = 1e3 * float(np.mean(times))
    fps = 1e3 / ms
    line = {
        "impl": "reference", "metric": "1080p frames/s at 1M Gaussians SH3 (DepthFirst mono frame)", "value": fps,
        "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": base_config(spec, fr.header.visibleCount, fr.header.totalInstances),
        "note": "CPU port of the reference's Metal kernels (oracle/), full frames",
        "cpu_baseline": cpu_baseline_dict(fps, cores, f"{args.steps} full frames of the workload", fr),
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)
    return 0


def e2e_sharded_upload(torch, dist, dev, rank, world, rs, pg, ph, N, K, cam, W, H, outs, steps, rec_bytes, sh_bytes):
    """End to end at N GPUs with every input byte crossing PCIe ONCE per step instead of N times: each rank
    uploads its 1/N slice of the step's scene from pinned host memory, the slices are replicated over NVLink with one NCCL
    all-gather per array (a plain copy collective: there is no compute to fuse it with), then every rank renders its view and
    downloads its frame. Two steps in flight on two streams / renderers / buffer sets. Returns seconds for `steps` steps."""
    chunk = (N + world - 1) // world
    lo, hi = min(rank * chunk, N), min((rank + 1) * chunk, N)
    st = [torch.cuda.Stream(device=dev) for _ in range(2)]
    dg = [torch.zeros(chunk * world * rec_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    dh = [torch.zeros(chunk * world * sh_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    sg = [torch.zeros(chunk * rec_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    sh_ = [torch.zeros(chunk * sh_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    dc = [torch.zeros((H, W, 4), dtype=torch.float16, device=dev) for _ in range(2)]
    dd = [torch.zeros((H, W), dtype=torch.float16, device=dev) for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]
    from gsm_renderer_b200.renderer import GaussianInput

    def step(i):
        j = i & 1
        if i >= 2:
            done[j].synchronize()
        with torch.cuda.stream(st[j]):
            sg[j][: (hi - lo) * rec_bytes].copy_(pg[lo * rec_bytes: hi * rec_bytes], non_blocking=True)
            sh_[j][: (hi - lo) * sh_bytes].copy_(ph[lo * sh_bytes: hi * sh_bytes], non_blocking=True)
            dist.all_gather_into_tensor(dg[j], sg[j])
            dist.all_gather_into_tensor(dh[j], sh_[j])
            rs[j].render(st[j], dc[j], dd[j], GaussianInput(dg[j], dh[j], N, K), cam, W, H)
            outs[j][0].copy_(dc[j], non_blocking=True)
            outs[j][1].copy_(dd[j], non_blocking=True)
            done[j].record(st[j])

    for i in range(4):
        step(i)
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        step(i)
    done[0].synchronize()
    done[1].synchronize()
    torch.cuda.synchronize()
    return time.perf_counter() - t0, int((hi - lo) * (rec_bytes + sh_bytes))


def bind_near_gpu(torch, local):
    """Pin this rank's process to the CPUs NVML reports as local to its GPU, BEFORE any pinned host buffer is allocated, so that the
    e2e path's staging memory is first touched on the GPU's NUMA node. Returns (note, previous affinity); any
    failure leaves the affinity alone."""
    prev = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    try:
        import pynvml as nv
        nv.nvmlInit()
        pr = torch.cuda.get_device_properties(local)
        bus = "%08x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        h = nv.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        near = {64 * w + b for w, v in enumerate(words) for b in range(64) if (int(v) >> b) & 1}
        want = near & prev if prev is not None else set()
        if want and want != prev:
            os.sched_setaffinity(0, want)
            return f"bound to {len(want)} of {len(prev)} CPUs local to GPU {bus}", prev
        return f"all {len(prev) if prev else 0} CPUs are local to GPU {bus} (or no overlap): not bound", prev
    except Exception as e:  # noqa: BLE001
        return f"not bound ({type(e).__name__})", prev


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between steps")
    ap.add_argument("--check", action="store_true", help="also verify the frame against the oracle (slow)")
    ap.add_argument("--no-modes", action="store_true", help="skip the multi-GPU shard modes (c3_strips, c5_views, c4_eyes)")
    ap.add_argument("--small-modes", action="store_true", help="run the modes on reduced clouds (smoke test of the plumbing)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from gsm_renderer_b200.renderer import (CameraParams, DepthFirstRenderer, GaussianColorSpace, GaussianInput,
                                            RendererConfig, RenderPrecision)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    affinity_note, prev_affinity = bind_near_gpu(torch, local)
    if world > 1:
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))

    cloud, g, h, spec = build_workload(args.workload)
    N, deg, prec, W, H, desc = spec
    K = syn.SH_COEFFS[deg]
    cfg = RendererConfig(maxGaussians=N, maxWidth=W, maxHeight=H,
                         precision=RenderPrecision.float16 if prec == "float16" else RenderPrecision.float32,
                         gaussianColorSpace=GaussianColorSpace.linear)  # PLYBenchmarkTests.swift:157-164
    r = DepthFirstRenderer(device=local, config=cfg)
    # views shard across ranks with no collective: rank k renders its own camera of a seeded orbit
    # every rank renders the C2 view: per-GPU work is fixed as N grows (weak scaling of an independent shard; different
    # poses would make `max over ranks` measure the heaviest view instead of the scaling)
    view, pos = np.eye(4, dtype=np.float32), np.zeros(3, np.float32)
    proj = syn.make_projection_matrix(W, H, NEAR, FAR)
    fx, fy = syn.focal_lengths(W, H)
    cam = CameraParams(view, proj, pos, fx, fy, NEAR, FAR)

    tg = torch.from_numpy(g.view(np.uint8).reshape(-1)).to(dev)
    th = torch.from_numpy(h.view(np.uint8).reshape(-1)).to(dev)
    color = torch.zeros((H, W, 4), dtype=torch.float16, device=dev)
    depth = torch.zeros((H, W), dtype=torch.float16, device=dev)
    inp = GaussianInput(tg, th, N, K)
    stream = torch.cuda.current_stream()
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step():
        r.render(stream, color, depth, inp, cam, W, H)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    t_w = time.perf_counter()
    extra = 0
    while time.perf_counter() - t_w < 0.3:  # let clocks settle under load before timing (untimed extra warm-up)
        step()
        extra += 1
    torch.cuda.synchronize()
    hd = r.debugReadHeader()
    V, I = hd.visibleCount, hd.totalInstances
    active = r.debugReadActiveTileCount()
    T = ((W + 15) // 16) * ((H + 15) // 16)
    max_per_tile = int(r.debugReadTileHeaders(T)[:, 1].max())

    # ---- timed region: per-step CUDA events on the launching stream, L2 flushed between steps
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    wall0 = time.perf_counter()
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        ev0[i].record(stream)
        step()
        ev1[i].record(stream)
    torch.cuda.synchronize()
    wall1 = time.perf_counter()
    wall = wall1 - wall0
    clocks = sampler.stop(wall0, wall1)
    step_ms = [ev0[i].elapsed_time(ev1[i]) for i in range(args.steps)]
    total_ms = float(sum(step_ms))
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
        dist.barrier()
    ms_per_step = total_ms / args.steps
    value = world * 1e3 / ms_per_step

    # ---- per-stage times: the same steps again with the C ABI's stage events on (untimed for `value`; the event
    # records between kernels add about a microsecond each, so the stages sum to slightly more than ms_per_step)
    r.setProfiling(True)
    stage_acc = {}
    for i in range(args.steps):
        if flush is not None:
            flush.fill_(i & 0xFF)
        step()
        torch.cuda.synchronize()
        for k, v in r.stageTimesMs().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    r.setProfiling(False)
    stage_ms = {k: v / args.steps for k, v in stage_acc.items()}

    # ---- e2e: host buffers through gsm_render_host (H2D + frame + D2H every step)
    pg = torch.from_numpy(g.view(np.uint8).reshape(-1)).pin_memory()
    ph = torch.from_numpy(h.view(np.uint8).reshape(-1)).pin_memory()
    pc = torch.zeros((H, W, 4), dtype=torch.float16).pin_memory()
    pd = torch.zeros((H, W), dtype=torch.float16).pin_memory()
    for _ in range(2):
        r.renderHost(pg, ph, N, K, cam, W, H, pc, pd)
    e2e_steps = max(4, min(args.steps, 20))
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        r.renderHost(pg, ph, N, K, cam, W, H, pc, pd)
    e2e_sync_s = time.perf_counter() - t0
    # two frames in flight (the reference's render() only encodes; apps double-buffer command buffers): a second
    # renderer with its own arena, stream and pinned outputs uploads frame i+1 while frame i renders and downloads
    r2 = DepthFirstRenderer(device=local, config=cfg)
    pc2 = torch.zeros((H, W, 4), dtype=torch.float16).pin_memory()
    pd2 = torch.zeros((H, W), dtype=torch.float16).pin_memory()
    rs, outs = (r, r2), ((pc, pd), (pc2, pd2))
    r2.renderHost(pg, ph, N, K, cam, W, H, pc2, pd2)
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        j = i & 1
        if i >= 2:
            rs[j].waitHost()
        rs[j].renderHostAsync(pg, ph, N, K, cam, W, H, outs[j][0], outs[j][1])
    rs[0].waitHost()
    rs[1].waitHost()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s, e2e_sync_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s, e2e_sync_s = float(t[0].item()), float(t[1].item())
    e2e_fps = world * e2e_steps / e2e_s
    e2e_sync_fps = world * e2e_steps / e2e_sync_s
    h2d = int(pg.numel() + ph.numel())
    d2h = int(pc.numel() * 2 + pd.numel() * 2)
    same = bool(torch.equal(pc.to(dev), color)) and bool(torch.equal(pc2.to(dev), color))
    e2e_extra = {"upload": "every rank uploads the whole scene" if world > 1 else "one GPU uploads the whole scene",
                 "h2d_bytes_per_step_per_rank": h2d, "h2d_bytes_per_step_all_ranks": h2d * world}
    if world > 1:
        # every rank needs the same scene: upload 1/N of it per rank over PCIe, replicate over NVLink (NCCL all-gather)
        rec_b, sh_b = (32 if prec == "float16" else 48), 3 * K * (2 if prec == "float16" else 4)
        sh_steps = max(8, min(args.steps, 40))
        sh_s, h2d_rank = e2e_sharded_upload(torch, dist, dev, rank, world, rs, pg, ph, N, K, cam, W, H, outs, sh_steps, rec_b, sh_b)
        t = torch.tensor([sh_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sh_s = float(t.item())
        same = same and bool(torch.equal(pc.to(dev), color)) and bool(torch.equal(pc2.to(dev), color))
        e2e_extra = {"upload": f"each rank uploads 1/{world} of the scene over PCIe, one NCCL all-gather per array replicates it over NVLink",
                     "h2d_bytes_per_step_per_rank": h2d_rank, "h2d_bytes_per_step_all_ranks": h2d, "steps_sharded": sh_steps,
                     "replicated_upload_value": e2e_fps}
        e2e_fps = world * sh_steps / sh_s

    # ---- device-resident throughput with two frames in flight (two renderers on two streams): what a multi-view batch
    # (config C5) gets per GPU -- one frame's latency-bound sorts run under the other's blend. Reported beside `value`
    # (which stays the single-frame number the reference's benchmark measures), never instead of it.
    s2 = torch.cuda.Stream(device=dev)
    color2 = torch.zeros_like(color)
    depth2 = torch.zeros_like(depth)
    n2 = max(8, 2 * (args.steps // 2))
    for _ in range(2):
        r.render(stream, color, depth, inp, cam, W, H)
        r2.render(s2, color2, depth2, inp, cam, W, H)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ea.record(stream)
    s2.wait_event(ea)
    for i in range(n2 // 2):
        r.render(stream, color, depth, inp, cam, W, H)
        r2.render(s2, color2, depth2, inp, cam, W, H)
    stream.wait_stream(s2)
    eb.record(stream)
    torch.cuda.synchronize()
    two_ms = ea.elapsed_time(eb) / n2
    if world > 1:
        t = torch.tensor([two_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        two_ms = float(t.item())
    two_in_flight = {"frames_per_s": world * 1e3 / two_ms, "ms_per_frame": two_ms, "frames": n2, "same_image": bool(torch.equal(color2, color)),
                     "note": "two renderers, two streams, no L2 flush between frames (frames overlap, so there is no 'between')"}
    del r2

    # ---- the multi-GPU shard modes (all ranks take part; at N=1 they give the single-GPU numbers the efficiencies refer to)
    modes = None
    if not args.no_modes:
        r.close()
        del r, tg, th, inp, color2, depth2, flush
        torch.cuda.empty_cache()
        import bench_modes
        modes = bench_modes.run_modes(torch, dist if world > 1 else None, rank, world, local, small=args.small_modes)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines
    peak, peak_src = measured_peaks()
    Vp = V  # Gaussians reaching the SH fetch >= V; the (9) exit after the fetch is rare. Lower bound used.
    sb = stage_bytes(N, V, Vp, I, T, W * H, 32 if prec == "float16" else 48, 3 * K * (2 if prec == "float16" else 4))
    stage_roofline = []
    # stage intervals with no kernel of their own are folded into the stage that does their work, bytes and time alike: the scan
    # runs inside the expansion kernel, the tile ranges are written by the tile sort's place kernel
    fold = {"expand": "applyScan", "tileSort": "ranges"}
    sb = dict(sb)
    sb["tileSort"] += sb.pop("ranges")
    for k in ("project", "depthSort", "expand", "tileSort", "blend"):
        ms = stage_ms.get(k, 0.0) + stage_ms.get(fold.get(k, ""), 0.0)
        gbs = sb[k] / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
        stage_roofline.append({"stage": k, "ms": ms, "bytes": int(sb[k]), "GBps": gbs, "frac": gbs / peak})
    def _profile_doc(name):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            return {}
    traffic_doc, pipes_doc = _profile_doc("traffic.json"), _profile_doc("pipes.json")
    # The dominant KERNEL, not the longest stage interval: a stage interval can hold several kernels (project = projection +
    # compaction, tileSort = three kernels). Each interval is split by its kernels' shares in the committed ncu capture.
    kern_us = pipes_doc.get("kernels_us") or []
    def _largest_kernel_share(stage):
        us = [k["us"] for k in kern_us if k.get("stage") == stage]
        return max(us) / sum(us) if us else 1.0
    dom = max(stage_roofline, key=lambda s: s["ms"] * _largest_kernel_share(s["stage"]))
    traffic, traffic_src = traffic_doc.get(dom["stage"]), traffic_doc.get("_source") or traffic_doc.get("note")
    pipes, pipes_src = pipes_doc.get(dom["stage"]), pipes_doc.get("_source") or pipes_doc.get("note")
    hbm = {"achieved": dom["GBps"], "peak": peak, "unit": "GB/s", "frac": dom["frac"], "peak_source": peak_src,
           "algorithmic_bytes": dom["bytes"], "ms": dom["ms"]}
    if dom["stage"] == "blend" and pipes and pipes.get("issue_active_pct"):
        # the blend is bound by instruction issue and the FMA / XU pipes, not by bytes: its roofline is the busiest of
        # those three from the ncu capture named in pipes_source; the live HBM figure stays beside it
        cands = {"issue": ("% of issue slots active (ncu smsp__issue_active)", pipes.get("issue_active_pct") or 0.0),
                 "fp_pipe": ("% of FMA-pipe cycles active (ncu sm__pipe_fma_cycles_active)", pipes.get("fma_pipe_cycles_active_pct") or 0.0),
                 "xu_pipe": ("% of XU-pipe issue (ncu sm__inst_executed_pipe_xu)", pipes.get("xu_pipe_inst_pct") or 0.0)}
        bound = max(cands, key=lambda k: cands[k][1])
        roofline = {"bound": bound, "kernel": "blend", "achieved": cands[bound][1], "peak": 100.0, "unit": cands[bound][0],
                    "frac": cands[bound][1] / 100.0, "traffic": traffic, "traffic_source": traffic_src, "pipes": pipes,
                    "pipes_source": pipes_src, "hbm": hbm}
    else:
        roofline = {"bound": "hbm", "kernel": dom["stage"], **hbm, "traffic": traffic, "traffic_source": traffic_src,
                    "pipes": pipes, "pipes_source": pipes_src}
    roofline["sort_stages_hbm_frac"] = {x["stage"]: x["frac"] for x in stage_roofline if x["stage"] in ("depthSort", "tileSort")}
    sort_blend_ms = stage_ms.get("tileSort", 0) + stage_ms.get("ranges", 0) + stage_ms.get("blend", 0)
    library_bar = library_sort_bar(torch, V, I, T, stage_ms)

    # ---- CPU baseline (bounded sample: whole frames of the same workload on this box's cores)
    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        if prev_affinity is not None:
            os.sched_setaffinity(0, prev_affinity)   # the CPU arm runs on every core the box gives this job
        run, fr, ob = oracle_frame_runner(g, h, spec)
        run()
        n_s, t_acc = 0, 0.0
        while n_s < 3 and t_acc < 20.0:
            t_acc += run()
            n_s += 1
        cpu_baseline = cpu_baseline_dict(n_s / t_acc, ob.lib().gsmo_num_threads(), f"{n_s} full frames of the workload after 1 warm-up", fr)
        if args.check:
            assert fr.header.visibleCount == V and fr.header.totalInstances == I

    line = {
        "metric": "1080p frames/s at 1M Gaussians SH3 (DepthFirst mono frame)", "value": value, "unit": "frames/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": base_config(spec, V, I),
        "run": {"activeTiles": active, "tiles": T, "maxInstancesPerTile": max_per_tile, "overflow": hd.overflow,
                "parallelism": f"views sharded over {world} GPU(s), one C2 view per GPU per step, no collective",
                "cpu_affinity": affinity_note,
                "l2": "inputs+arena > L2 and " + ("no flush" if args.no_flush else "L2 flushed between steps (256 MiB write, untimed)"),
                "timing": "per-step CUDA events on the launching stream, summed; max over ranks",
                "warmup_extra_steps": extra},
        "mtile_instances_per_s": (I / (sort_blend_ms * 1e-3) / 1e6) if sort_blend_ms > 0 else None,
        "two_frames_in_flight": two_in_flight,
        "step_ms": {"min": float(min(step_ms)), "median": float(np.median(step_ms)), "max": float(max(step_ms))},
        "stage_ms": stage_ms,
        "roofline": roofline,
        "stage_roofline": stage_roofline,
        "library_bar": library_bar,
        "cpu_baseline": cpu_baseline,
        "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h * world,
                "steps": e2e_steps, "matches_device_path": same, "frames_in_flight": 2,
                "one_frame_in_flight": e2e_sync_fps, **e2e_extra,
                "note": "pinned host buffers; every step uploads the scene and downloads colour+depth of every rank's frame, two "
                        "steps in flight; N=1: gsm_render_host_async/_wait (one_frame_in_flight = blocking gsm_render_host)"},
        "modes": modes,
        "gpu_launches": KERNELS_PER_FRAME * args.steps,
        "clocks": clocks,
        "wall_s_timed_region": wall,
    }
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
