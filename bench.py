#!/usr/bin/env python
"""Headline benchmark: voxels/s for forward+backward of Res16UNet34C on synthetic 200k-voxel
ScanNet-shaped scenes (BASELINE.json metric, configs[1]) on 1..8 B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one scene (batch 1) through the whole hot path: coordinate-manager construction from the
[N,4] coordinate matrix (hash insert, 4 stride maps, 9 kernel maps), backbone forward, scalar loss,
backward through every layer.  Scenes are independent, so N GPUs run N scenes per step (weak scaling);
at N > 1 the step is the data-parallel training step of BASELINE configs[4]: the ~151 MB of fp32 gradients
are averaged over the ranks by bucketed NCCL all-reduces launched from gradient hooks while backward is
still running (unscene3d_b200/distributed.py).  The timed region is bracketed by barrier + synchronize,
timed with CUDA events, and the maximum over ranks is reported by rank 0 as ONE JSON line.

Legs:
  value      inputs resident in HBM, device-timed, all ranks
  e2e        same step fed from pinned HOST buffers (H2D inside the timed region) + D2H of the loss
  roofline   per-launch CUDA-event timing of the dominant kernel (sparse-conv gather, fwd + dgrad)
             against its algorithmic bytes and MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (oracle/me_cpu.py, "port": same gather-GEMM-scatter algorithm as
             MinkowskiEngine's CPU path, which is not installable here) on the host cores, rank 0, N=1
  --impl reference   the same CPU oracle as the reference arm, on the SAME 200k-voxel scene and seed (see DESIGN.md:
             MinkowskiEngine is absent, so the reference's CPU path is its restatement, kind "port")
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
sys.path.insert(0, ROOT)
# growing a caching-allocator pool with cudaMalloc synchronises the device; mapped (expandable) segments do not
os.environ.setdefault("PYTORCH_CUDA_ALLOC_CONF", "expandable_segments:True")

METRIC = "voxels/sec fwd+bwd Res16UNet34C @200k-voxel scenes"
UNIT = "voxels/s"
N_VOXELS = 200_000
# dram__bytes_read.sum + dram__bytes_write.sum of the largest launch of each kernel, from the ncu --set full capture summarised
# in profiles/ (see there for the command), next to the algorithmic bytes of that launch
NCU_TRAFFIC = {"mt": {"launch": "200k voxels, k3, 128 -> 96 (pattern order)", "dram_bytes": 497897984, "alg_bytes": 180500000,
                      "l2_to_sm_bytes": 2117809000, "duration_us": 280.3,  # the launch is bound by the L2 -> SM path: 7.6 TB/s
                      "source": "profiles/r2_final_ncu_full_summary.md"},
               "wgrad": {"launch": "200k voxels, k3, 128 -> 96", "dram_bytes": 225198336, "alg_bytes": 180500000,
                         "l2_to_sm_bytes": 1905437000, "duration_us": 362.0,
                         "source": "profiles/r2_final_ncu_full_summary.md"}}
DTYPE = "bf16x3"  # tcgen05 bf16 products, three-term split hi*hi + lo*hi + hi*lo (fp32-faithful), fp32 accumulation in TMEM
PRIME_STEPS = 30  # untimed steps BEFORE the --warmup steps: allocator pools of the two streams, NVML, clock / power ramp


def bench_config(voxels, world):
    """`config` of the JSON line — identical for both arms (the reference arm runs the same workload on the host cores)."""
    return {"workload": f"Res16UNet34C fwd+bwd, synthetic ScanNet-shaped {voxels}-voxel scene (seed = rank), batch=1 per GPU (BASELINE configs[1])",
            "parallelism": "dp1 (one scene per GPU)" if world == 1 else
                           f"dp{world}: one scene per GPU, gradients averaged by bucketed NCCL all-reduce overlapped with backward (BASELINE configs[4])",
            "l2": "256 MiB buffer written between timed GPU iterations (L2 flush)",
            "step": "coordinate-manager build (on a high-priority side stream) + forward + loss + backward"
                    + ("" if world == 1 else " + gradient all-reduce")}


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons during the timed region: in-process NVML queries (nvidia_ml_py) issued by the timed
    loop itself; `nvidia-smi -lms` in a subprocess as a fallback.  (Measured on the B200 box: NVML / nvidia-smi queries
    stall kernel launches — a looping nvidia-smi slows this host-bound step from ~20 to 28 ms, a query per step to 24 ms —
    so only two samples are taken, from the loop thread.)"""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index, period_s=0.02):
        self.gpu, self.samples, self.proc, self.period = gpu_index, [], None, period_s
        self.nvml, self.handle, self.stop_flag, self.thread, self.query_ms = None, None, False, None, []

    def start(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES remaps indices: resolve through the PCI bus id of the torch device
            import torch

            bus = torch.cuda.get_device_properties(self.gpu).pci_bus_id if hasattr(torch.cuda.get_device_properties(self.gpu), "pci_bus_id") else None
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.gpu) if bus is None else self._handle_by_bus(pynvml, bus)
            self.nvml = pynvml
            self.smax = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            return  # sampled from the timed loop itself (sample()), one NVML query per step: no second thread, no GIL traffic
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    @staticmethod
    def _handle_by_bus(pynvml, bus_id):
        for i in range(pynvml.nvmlDeviceGetCount()):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            if int(pynvml.nvmlDeviceGetPciInfo(h).bus) == int(bus_id):
                return h
        return pynvml.nvmlDeviceGetHandleByIndex(0)

    def sample(self):
        """One in-process NVML query (a few tens of microseconds); called once per step inside the timed region."""
        nv = self.nvml
        if nv is None:
            return
        try:
            tq = time.time()
            sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
            reasons = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            self.samples.append((time.time(), sm, reasons))
            self.query_ms.append((time.time() - tq) * 1e3)
        except Exception:
            pass

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.nvml is not None:
            nv = self.nvml
            names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            sm, reasons = [], set()
            for ts, clk, bits in self.samples:
                if t0 - 0.02 <= ts <= t1 + 0.02:
                    sm.append(clk)
                    reasons |= {k for k, v in names.items() if bits & v}
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.smax, "reasons": sorted(reasons), "samples": len(sm),
                    "source": "nvml", "query_ms": [round(q, 2) for q in self.query_ms]}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        for ts, line in self.samples:
            if not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax = float(p[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm),
                "source": "nvidia-smi"}


# ------------------------------------------------------------------------------------------- CPU arm
def cpu_oracle_run(steps, warmup, n_voxels, seed=0, small_warmup=True):
    """Times the CPU oracle (Res16UNet34C fwd+bwd incl. coordinate-manager construction) on all host threads."""
    import torch

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import Cfg, our_models_on_oracle
    from oracle import me_cpu
    from unscene3d_b200_synthetic import make_scene

    torch.set_num_threads(os.cpu_count() or 1)
    net = our_models_on_oracle().res16unet.Res16UNet34C(3, 20, Cfg(), D=3, out_fpn=True).train()
    scene = make_scene(n_voxels, seed=seed, with_masks=False)
    c4 = torch.from_numpy(np.concatenate([np.zeros((scene.n, 1), np.int32), scene.coords], 1))
    feats = torch.from_numpy(scene.colors)
    w = torch.linspace(-1, 1, 96)

    def step():
        x = me_cpu.SparseTensor(feats, c4)
        out, _ = net(x)
        (out.F * w).mean().backward()
        net.zero_grad(set_to_none=True)

    if small_warmup:  # cpu_baseline leg of the GPU arm: bounded — warm the thread pool / allocator on a 5k-voxel scene
        small = make_scene(5000, seed=seed + 1, with_masks=False)
        c4s = torch.from_numpy(np.concatenate([np.zeros((small.n, 1), np.int32), small.coords], 1))
        for _ in range(max(warmup, 1)):
            out, _ = net(me_cpu.SparseTensor(torch.from_numpy(small.colors), c4s))
            (out.F * w).mean().backward()
            net.zero_grad(set_to_none=True)
    else:
        for _ in range(warmup):
            step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return n_voxels * steps / dt, dt / steps * 1e3, torch.get_num_threads()


def _load_synthetic_standalone():
    """The scene generator is pure numpy; load it without importing the CUDA package (CPU arm)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location("unscene3d_b200_synthetic", os.path.join(ROOT, "unscene3d_b200", "synthetic.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["unscene3d_b200_synthetic"] = mod
    spec.loader.exec_module(mod)
    return mod


def run_reference(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0
    _load_synthetic_standalone()
    # same workload as the GPU arm: the full scene, same generator and seed; one step is ~4 s on 16 host threads
    v, ms, threads = cpu_oracle_run(args.steps, args.warmup, args.voxels, seed=0, small_warmup=False)
    world = env_int("WORLD_SIZE", 1)
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(args.voxels, 1 if world == 1 else world),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{args.steps} steps of the full {args.voxels}-voxel scene (seed 0) after {args.warmup} warm-up steps; CPU oracle = "
                                   "gather-GEMM-scatter restatement of the MinkowskiEngine CPU path (ME is not installable offline)"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


# ------------------------------------------------------------------------------------------- GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    import unscene3d_b200
    from unscene3d_b200 import _lib, engine, models
    from unscene3d_b200 import distributed as D
    from unscene3d_b200.engine import functional as Fn
    from unscene3d_b200.synthetic import level_sizes, make_scene
    from unscene3d_b200.utils import BackboneConfig, conv_layer_bytes, seeded_state

    scene = make_scene(args.voxels, seed=rank, with_masks=False)
    c4_host = torch.from_numpy(np.concatenate([np.zeros((scene.n, 1), np.int32), scene.coords], 1)).pin_memory()
    f_host = torch.from_numpy(scene.colors).pin_memory()
    c4_dev, f_dev = c4_host.to(dev), f_host.to(dev)

    net = models.Res16UNet34C(3, 20, BackboneConfig(), D=3, out_fpn=True)
    net.load_state_dict(seeded_state(net, 0))
    net = net.to(dev).train()
    wvec = torch.linspace(-1, 1, 96, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    loss_host = torch.zeros(1).pin_memory()

    # Data-parallel half of the step (N > 1): ~151 MB of fp32 gradients in 32 MB buckets, each all-reduced (NCCL, AVG) as soon
    # as backward has produced its last gradient; finish() makes the compute stream wait for the collectives.
    reducer = D.GradientReducer(net.parameters(), bucket_bytes=32 << 20) if world > 1 else None
    comm_on = {"on": True}
    params_ = list(net.parameters())

    def step(coords, feats):
        # a training step packs the (updated) weights of every convolution once — all images in one launch, as a loop does
        # after its optimizer update; with no optimizer in the metric the cached images would otherwise survive from step
        # to step
        Fn.pack_network(net)
        x = engine.SparseTensor(feats, coords)
        out, _ = net(x)
        loss = (out.F * wvec).mean()
        loss.backward()
        if reducer is not None:
            reducer.enabled = comm_on["on"]
            reducer.finish()
            reducer.zero_grad()
        else:
            for p_ in params_:  # == net.zero_grad(set_to_none=True) without walking the module tree every step
                p_.grad = None
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, sampler=None):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        s.record()
        # NVML queries contend with kernel launches for the driver lock (measured: one query per step costs 4.5 ms/step at
        # one GPU and 12 ms/step at two, on this launch-heavy step).  The clocks are therefore sampled once at the middle
        # of the timed steps and twice more right after the closing event has been queued, while the device is still
        # working through the launches the host has queued ahead — inside the timed region, off its critical path.
        host = []
        for i in range(steps):
            th = time.perf_counter()
            flush.zero_()
            fn()
            if sampler is not None and i + 1 == max(steps // 2, 1):
                sampler.sample()
            host.append((time.perf_counter() - th) * 1e3)
        e.record()
        host_ms[fn.__name__] = host
        if sampler is not None:
            sampler.sample()
            sampler.sample()
        barrier()
        t1 = time.time()
        ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), t0, t1

    host_ms = {}  # host-side duration of every timed step (queueing only): an outlier here is a host stall, not device time

    def step_resident():
        step(c4_dev, f_dev)

    # Coordinate maps are built on a dedicated high-priority stream (engine.set_coordinate_stream): their size read-back
    # then waits for the map's own integer kernels only, not for the previous step's backward pass on the compute stream.
    # The end-to-end leg issues the host-to-device copies of a step on the same stream, as a prefetching data loader does.
    side = torch.cuda.Stream(device=dev, priority=-1)
    engine.set_coordinate_stream(side, dev)

    def step_e2e():
        main = torch.cuda.current_stream(dev)
        with torch.cuda.stream(side):
            c = c4_host.to(dev, non_blocking=True)
            f = f_host.to(dev, non_blocking=True)
        main.wait_stream(side)  # the features are consumed on the compute stream
        c.record_stream(main)
        f.record_stream(main)
        loss = step(c, f)
        loss_host.copy_(loss.detach().reshape(1), non_blocking=True)  # pinned target; the timed region ends with a synchronize

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # The caching allocator needs a handful of steps to stop growing its pools (two streams allocate: compute and
    # coordinate); a cudaMalloc inside the timed region synchronises the device.  Measured: 3 warm-up steps leave the first
    # timed leg at 22.8 ms/step, the same leg after ~15 steps runs at 19.6; one second of steps (50) also rides out the
    # clock / power ramp of a fresh box (one 2-GPU run with 20 warm-up steps measured 20.7 ms, three repeats 19.04).
    if world > 1:
        # first collectives of the communicator (buffer allocation, proxy threads) well before the timed legs: at 8 ranks the
        # first timed leg measured 21.3 ms/step against 19.3 for the identical second leg
        for _ in range(3):
            barrier()
            dist.all_reduce(torch.zeros(1, device=dev))
    # PRIME_STEPS untimed steps (allocator pools of the two streams, NVML's lazy initialisation, clock / power ramp of a fresh
    # box; lowered only for ncu launch lists) — then everything alive is parked in the permanent GC generation (a full
    # cyclic collection over the whole heap inside a timed step is a ~100 ms host stall; training loops do the same with
    # gc.freeze()) — then exactly --warmup warm-up steps, then the clock.
    import gc

    n_prime = int(os.environ.get("US3D_BENCH_PRIME", str(PRIME_STEPS)))
    debug = bool(os.environ.get("US3D_BENCH_DEBUG"))
    if debug:
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ.get("US3D_BENCH_DUMP_AFTER", "90")), exit=True)
    for i in range(n_prime):
        step_resident()
        if debug:
            torch.cuda.synchronize()
            print(f"[debug] rank {rank}: prime step {i} done", file=sys.stderr, flush=True)
        if rank == 0:
            sampler.sample()  # first query: 3 ms; samples outside the timed region are dropped
    if rank == 0 and sampler.nvml is None:
        time.sleep(0.3)  # nvidia-smi fallback: let the subprocess start reporting
    gc.collect()
    gc.freeze()
    for _ in range(args.warmup):
        step_resident()
    _lib.reset_launch_count()
    ms_total, t0, t1 = timed(step_resident, args.steps, sampler if rank == 0 else None)
    launches = _lib.launch_count()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    for _ in range(max(args.warmup, 5)):  # the end-to-end leg has its own allocator pool (copies on the coordinate stream)
        step_e2e()
    gc.collect()
    gc.freeze()
    ms_e2e, _, _ = timed(step_e2e, args.steps)

    # ---- N > 1: what the collective costs.  (a) the same step with the all-reduces switched off (the buckets are still
    # filled): the difference is the EXPOSED communication per step; (b) the buckets all-reduced back to back with nothing
    # else running: algorithmic bus bandwidth 2 (N-1)/N * bytes / time.
    comm = None
    if reducer is not None:
        comm_on["on"] = False
        for _ in range(3):
            step_resident()
        ms_nocomm, _, _ = timed(step_resident, args.steps)
        comm_on["on"] = True
        for _ in range(2):
            step_resident()
        barrier()
        s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        s_ev.record()
        for _ in range(reps):
            for bkt in reducer.buckets:
                dist.all_reduce(bkt.flat, op=dist.ReduceOp.AVG)
        e_ev.record()
        torch.cuda.synchronize()
        ar_ms = torch.tensor([s_ev.elapsed_time(e_ev) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(ar_ms, op=dist.ReduceOp.MAX)
        ar_ms = float(ar_ms.item())
        nbytes = reducer.total_bytes
        comm = {"collective": "all-reduce (NCCL, AVG) of the fp32 gradients, launched per bucket from gradient hooks during backward",
                "bytes_per_step": int(nbytes), "buckets": len(reducer.buckets),
                "step_ms_without_collective": ms_nocomm / args.steps,
                "exposed_ms_per_step": (ms_total - ms_nocomm) / args.steps,
                "standalone_allreduce_ms": ar_ms,
                "busbw_gbs": 2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9}

    if os.environ.get("US3D_BENCH_DEBUG"):
        for name, fn in (("resident", step_resident), ("e2e", step_e2e), ("resident", step_resident), ("e2e", step_e2e)):
            ms_dbg, _, _ = timed(fn, args.steps)
            if rank == 0:
                print(f"[debug] {name}: {ms_dbg / args.steps:.2f} ms/step", file=sys.stderr, flush=True)
    value = world * args.voxels * args.steps / (ms_total * 1e-3)
    e2e_value = world * args.voxels * args.steps / (ms_e2e * 1e-3)

    roofline = roofline_wgrad = cpu_baseline = None
    if rank == 0:
        # ---- roofline leg.  The dominant kernel is us3d::mt::k_spconv_mt (tcgen05 sparse-conv forward + input gradient);
        # the second is us3d::wg::k_wgrad (weight gradient).  north_star judges the HBM roof:
        #   achieved = ALGORITHMIC bytes of the launches (fp32 X + Y [+ dX] rows + weights, kernel-map
        #   indices and BN / ReLU / residual traffic count zero) / their summed durations, CUDA events recorded inside libus3d
        #   around each launch on the launch stream during normal steps;  peak = MEASURED_PEAKS.json hbm_gbs.
        # The tensor view of the same launches (algorithmic FLOPs = 2 * kernel-map pairs * Cin * Cout against the measured
        # sustained bf16 GEMM rate) is kept next to it: at Cin, Cout >= 96 the layers are tensor-bound at ideal traffic
        # (264 FLOP/B against a ridge of 212), and the kernel executes ~7x the algorithmic FLOPs (three bf16
        # products per fp32 product, zero rows of 128-row tiles).
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peaks = json.load(open(peaks_path))
            peak_tf = float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
            peak_src = "measured (MEASURED_PEAKS.json: hbm_gbs; bf16_tflops_sustained for the tensor view — kernels timed inside a long step)"
            peak_hbm = float(peaks["hbm_gbs"])
        else:
            peak_tf, peak_src, peak_hbm = 1400.0, "fallback (B200_PROFILING.md: 6.65 TB/s, ~1.4 PFLOP/s sustained)", 6650.0
        prof_steps = 2
        recs = []
        comm_on["on"] = False  # rank 0 runs these steps alone: no collective may be issued here
        for _ in range(prof_steps):
            with Fn.KernelTimer() as kt:
                flush.zero_()
                step_resident()
            recs += kt.summary()
        # (kind, n_in, n_out, kvol, cin, cout, pairs, path, ms)
        passes = 3 if Fn.get_precision() == 3 else 1

        def view(rows, kinds_label, traffic):
            ms = sum(r[8] for r in rows)
            flops = sum(2.0 * r[6] * r[4] * r[5] for r in rows)
            exec_flops = sum(2.0 * r[2] * r[3] * r[4] * r[5] * passes for r in rows)
            nbytes = sum(conv_layer_bytes(r[1], r[2], r[3], r[4], r[5], r[0]) for r in rows)
            gbs = nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0
            tfs = flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0
            return {"bound": "hbm", "kernel": kinds_label, "achieved": gbs, "peak": peak_hbm, "unit": "GB/s", "frac": gbs / peak_hbm,
                    "traffic": traffic["dram_bytes"], "traffic_of": traffic, "peak_source": peak_src,
                    "timing": "CUDA events recorded around each launch inside libus3d, on the launch stream, during normal steps",
                    "launches_per_step": len(rows) // prof_steps, "kernel_ms_per_step": ms / prof_steps,
                    "alg_bytes_per_step": nbytes / prof_steps, "alg_gflop_per_step": flops / prof_steps / 1e9,
                    "tensor_view": {"achieved_tflops": tfs, "peak_tflops": peak_tf, "frac": tfs / peak_tf,
                                    "executed_tflops": exec_flops / (ms * 1e-3) / 1e12 if ms > 0 else 0.0}}

        dom = [r for r in recs if r[0] in ("fwd", "dgrad") and r[7] == "mt"]
        wg = [r for r in recs if r[0] == "wgrad" and r[7] == "wgrad-tc"]
        # `traffic`: dram__bytes_read.sum + dram__bytes_write.sum of ONE launch (ncu --set full, profiles/r2_final_ncu_full_summary.md)
        # next to that launch's algorithmic bytes
        roofline = view(dom, "us3d::mt::k_spconv_mt (tcgen05 sparse-conv forward + input-gradient launches)", NCU_TRAFFIC["mt"])
        roofline_wgrad = view(wg, "us3d::wg::k_wgrad (tcgen05 weight-gradient launches)", NCU_TRAFFIC["wgrad"])
        other_ms = sum(r[8] for r in recs if r not in dom and r not in wg)
        roofline.update({"other_conv_ms_per_step": other_ms / prof_steps, "step_ms": ms_total / args.steps,
                         "level_sizes": level_sizes(c4_host.numpy())})
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, "gpurun_out", "bench_layers.json"), "w") as fh:
            json.dump([{"kind": r[0], "n_in": r[1], "n_out": r[2], "kvol": r[3], "cin": r[4], "cout": r[5], "pairs": r[6], "path": r[7],
                        "ms": r[8], "alg_bytes": conv_layer_bytes(r[1], r[2], r[3], r[4], r[5], r[0])} for r in recs[: len(recs) // prof_steps]], fh)
        # ---- CPU baseline leg (bounded sample), N=1 only
        if world == 1 and not args.no_cpu_baseline:
            _load_synthetic_standalone()
            v, ms, threads = cpu_oracle_run(1, 1, args.voxels)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                            "sample": f"1 step of the full {args.voxels}-voxel scene after a 5k-voxel warm-up ({ms:.0f} ms); CPU oracle = "
                                      "gather-GEMM-scatter restatement of the MinkowskiEngine CPU path (ME itself is not installable offline)"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE, "data": "synthetic",
            "config": bench_config(args.voxels, world),
            "prime_steps": n_prime,
            "host_ms_per_step": {k: {"median": float(np.median(v)), "max": float(np.max(v)), "argmax": int(np.argmax(v))}
                                 for k, v in host_ms.items()},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(c4_host.numel() * 4 + f_host.numel() * 4),
                    "d2h_bytes_per_step": 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "roofline_wgrad": roofline_wgrad,
            "cpu_baseline": cpu_baseline, "comm": comm,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--voxels", type=int, default=N_VOXELS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    world = env_int("WORLD_SIZE", 1)
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 1000), os.path.abspath(__file__)] + sys.argv[1:]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
