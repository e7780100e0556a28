#!/usr/bin/env python
"""bench.py -- advance_mu_t grid-points/s on N B200s of one node (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload conus3|conus12|weak2048|deep120|tiny]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference's own CPU C code on the host cores

A "step" is one pass of the hot path over the workload: for the default workload (conus3, the
CONUS-3km-class 1800x1060x50 grid BASELINE.json's target is quoted on) that is the device-resident
6-acoustic-step loop of one RK3 sub-step, i.e. six advance_mu_t calls over the whole grid, with the
u/v one-cell halo exchanged before every call when N > 1 (2-D patch decomposition, strong scaling).
Prints ONE JSON line (rank 0).  Every number is measured in this run; nothing is cached.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# name -> (nx, ny, nz, small steps per bench step, dx [m], dts [s], scaling, description)
WORKLOADS = {
    "conus3": (1800, 1060, 50, 6, 3000.0, 3.0, "strong",
               "CONUS 3 km-class grid 1800x1060x50, device-resident 6-acoustic-step loop"),
    "conus12": (425, 300, 35, 1, 12000.0, 12.0, "strong", "CONUS 12 km-class grid 425x300x35, single small step"),
    "weak2048": (2048, 2048, 80, 6, 3000.0, 3.0, "weak", "2048x2048x80 per-GPU tile, 6-acoustic-step loop"),
    "deep120": (512, 512, 120, 6, 3000.0, 3.0, "strong", "deep-column 512x512x120, 6-acoustic-step loop"),
    "tiny": (74, 61, 28, 1, 12000.0, 12.0, "strong", "driver-equivalent tiny domain 74x61x28"),
    "patch8": (1800, 133, 50, 6, 3000.0, 3.0, "strong", "one rank's j-slab of conus3 at 8 GPUs (tuning aid)"),
}
HALO = 5
EPSSM = 0.1
METRIC = "advance_mu_t grid-points/s"
UNIT = "grid-points/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                 "-i", str(self.gpu)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        try:
            for line in open(self.path):
                c = [x.strip() for x in line.split(",")]
                if len(c) < 9:
                    continue
                try:
                    sm.append(float(c[1])); smax.append(float(c[2])); power.append(float(c[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def make_global_grid(workload, n_gpus):
    import wrf_model_cuda_sample_b200 as wrf
    from wrf_model_cuda_sample_b200 import parallel
    nx, ny, nz, nsmall, dx, dts, scaling, desc = WORKLOADS[workload]
    if scaling == "weak" and n_gpus > 1:
        px, py = parallel.choose_process_grid(n_gpus, nx, ny * n_gpus)
        ny = ny * n_gpus          # per-GPU tile stays nx x (ny/N) = the named tile
    g = wrf.Grid.from_shape(nx, ny, nz, halo=HALO, periodic_x=False, specified=True, nested=False)
    scalars = (np.float32(1.0 / dx), np.float32(1.0 / dx), np.float32(dts), np.float32(EPSSM))
    return g, scalars, nsmall, dx


# -------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own C (oracle/_ref) on the host cores
# -------------------------------------------------------------------------------------------------
def cpu_reference_run(g, scalars, dx, steps, warmup, budget_s=25.0):
    """Times ONE advance_mu_t small step of the reference C over the full grid per step (a bounded
    sample of the workload step), j-tiled over all host threads as WRF tiles it.  -> dict."""
    import wrf_model_cuda_sample_b200 as wrf
    from oracle import loader
    cores = os.cpu_count() or 1
    # all the host threads: torchrun exports OMP_NUM_THREADS=1 to its children, which would silently turn
    # this into a single-thread measurement
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        import ctypes
        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(cores)
    except OSError:
        pass
    if loader.have_ref():
        fn, kind, what = loader.reference_c, "reference", "reference advance_mu_t.c (unmodified, gcc -O3 -ffp-contract=off)"
    else:
        fn, kind, what = loader.oracle_c, "port", "oracle/advance_mu_t_oracle.c (gcc -O3 -ffp-contract=off)"
    # bound the sample: shrink in j until one call is affordable (full grid normally fits easily)
    n3_full, _ = g.updated_points()
    rows = g.jde
    est = n3_full / (60e6 * max(1, cores) * 0.6)
    while est * (steps + warmup) > budget_s and rows > 64:
        rows //= 2
        est /= 2
    gs = g if rows == g.jde else wrf.Grid.from_shape(g.ide, rows, g.kde, halo=HALO, periodic_x=g.periodic_x,
                                                      specified=g.specified, nested=g.nested)
    f = wrf.synth_fields(gs, dx_m=dx)
    f.pop("__pinned__", None)
    n3, n2 = gs.updated_points()
    tiles = max(1, cores * 4)
    times = []
    for s in range(warmup + steps):
        t0 = time.perf_counter()
        fn(f, gs, scalars, tiles=tiles)
        dt = time.perf_counter() - t0
        if s >= warmup:
            times.append(dt)
    total = sum(times)
    return {"value": n3 * len(times) / total, "unit": UNIT, "cores": cores, "kind": kind,
            "sample": f"{len(times)} x one small step over {gs.ide}x{gs.jde}x{gs.kde} ({n3} points), {what}, "
                      f"OpenMP over {tiles} j-tiles on {cores} threads",
            "ms_per_small_step": 1e3 * total / len(times)}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    g, scalars, nsmall, dx = make_global_grid(args.workload, 1)
    steps = max(1, min(args.steps, 20))        # each step is one small step over the full grid: bounded sample
    warmup = max(1, min(args.warmup, 3))
    r = cpu_reference_run(g, scalars, dx, steps, warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_small_step"] * nsmall,
        "higher_is_better": True, "scaling": WORKLOADS[args.workload][6], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][7]}",
                   "grid": f"{g.ide}x{g.jde}x{g.kde}", "small_steps_per_step": nsmall,
                   "note": "each timed step is ONE small step (bounded sample); ms_per_step is scaled to the workload step"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import wrf_model_cuda_sample_b200 as wrf
    from wrf_model_cuda_sample_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: native libraries (NCCL prints its version banner to stdout)
    # are pointed at stderr for the whole run, the line itself is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1 and args.gpus == 1, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    g, scalars, nsmall, dx = make_global_grid(args.workload, world)
    if args.pgrid:
        px, py = (int(x) for x in args.pgrid.lower().split("x"))
    else:
        px, py = parallel.choose_process_grid(world, g.ide, g.jde)
    assert px * py == world
    decomp = parallel.Decomposition(g, px, py, halo=HALO)
    pg = decomp.patch_grid(rank)
    kernel = {"auto": wrf.KERNEL_AUTO, "pipe": wrf.KERNEL_PIPE, "tile": wrf.KERNEL_TILE, "column": wrf.KERNEL_COLUMN}[args.kernel]

    # ---- inputs: pinned host arrays (also the e2e source), uploaded once for the resident loop ----
    host = wrf.synth_fields(pg, pinned=True, dx_m=dx)
    pinned_keep = host.pop("__pinned__")
    patch = wrf.Patch(pg, device=local_rank)
    main = torch.cuda.Stream(device=dev)                       # a real stream: capturable, and what events time
    torch.cuda.set_stream(main)
    patch.set_stream(main.cuda_stream)
    patch.set_scalars(*scalars)
    patch.set_kernel(kernel)
    patch.upload(host)
    patch.sync()

    n3_local, n2_local = pg.updated_points()
    n3_global, n2_global = g.updated_points()
    bytes_local = pg.algorithmic_bytes()

    # ---- one bench step ----
    if world == 1:
        def step():
            patch.step_graph(nsmall)
        launches_per_step = nsmall
    else:
        halo = parallel.GpuPatchHalo(patch, decomp, rank, dev)
        ex = parallel.HaloExchanger(decomp, rank, halo.pack, halo.recv_buffer, halo.unpack)
        ex.exchange(parallel.CONSTANT_HALOS)                  # once per RK sub-step
        interior, strips = decomp.interior_and_boundary_tiles(rank)
        comm = torch.cuda.Stream(device=dev)
        n_p2p = len(ex.plan(parallel.STEP_HALOS))              # one pack or unpack kernel of ours per message
        launches_per_step = nsmall * ((1 if interior else 0) + len(strips) + n_p2p)

        def step():
            for _ in range(nsmall):
                comm.wait_stream(main)                         # u,v of the previous step are final
                with torch.cuda.stream(comm):
                    patch.set_stream(comm.cuda_stream)
                    tok = ex.start(parallel.STEP_HALOS)        # pack + NCCL send/recv on the comm stream
                    ex.finish(tok)                             # unpack into the halo cells
                    for s in strips:                           # the columns that read the received halo follow
                        patch.step(pg.with_tile(*s))           # on the SAME stream: they fill the SM slots the
                patch.set_stream(main.cuda_stream)             # interior kernel's last partial wave leaves idle
                if interior:
                    patch.step(pg.with_tile(*interior))        # overlaps exchange + strips
                main.wait_stream(comm)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    flush = None
    if args.workload in ("conus12", "tiny"):
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    for _ in range(args.warmup):
        step()
    barrier()

    # N > 1: the step is ~10 short launches per acoustic step (pack, NCCL send/recv, unpack, interior, strips)
    # driven from Python; capture the whole multi-stream step once in a CUDA graph so the timed region is
    # launch-bound on the GPU, not on the interpreter.  Falls back to eager launches if capture fails.
    exec_mode = "cuda graph (wrfb200_step_graph)" if world == 1 else "eager launches"
    if world > 1 and not args.no_graph:
        ok = torch.tensor([1], device=dev)
        try:
            eager_step = step
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=main, capture_error_mode="thread_local"):
                eager_step()
            patch.set_stream(main.cuda_stream)
            graph.replay()
            torch.cuda.synchronize()
        except Exception as e:                                  # pragma: no cover - depends on NCCL / driver
            ok.zero_()
            sys.stderr.write(f"[bench] rank {rank}: CUDA-graph capture failed ({type(e).__name__}: {e}); eager\n")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            step = graph.replay
            exec_mode = "cuda graph (torch.cuda.graph over pack / NCCL send-recv / unpack / interior / strips)"
        else:
            step = eager_step
        barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = patch.launch_count()
    if flush is None:
        # inputs are far larger than L2: time the K steps back to back
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(main)
        for _ in range(args.steps):
            step()
        e1.record(main)
        barrier()
        elapsed_ms = e0.elapsed_time(e1)
        l2_note = "inputs larger than L2 (%.2f GB touched per small step vs 126 MB L2)" % (bytes_local / 1e9)
    else:
        elapsed_ms = 0.0
        for _ in range(args.steps):
            flush.fill_(1)                                     # evict the previous step's lines from L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            e0.record(main)
            step()
            e1.record(main)
            barrier()
            elapsed_ms += e0.elapsed_time(e1)
        l2_note = "L2 flushed (256 MB write) before every timed step"
    launches = patch.launch_count() - l0
    if launches == 0:                                          # graph replays are not seen by the handle's counter
        launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([elapsed_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed_ms = float(t.item())
    ms_per_step = elapsed_ms / args.steps
    value = n3_global * nsmall * args.steps / (elapsed_ms * 1e-3)

    # ---- roofline of the dominant kernel (the tile kernel over this rank's patch) ----
    peak, peak_src = peaks()
    kernel_ms = elapsed_ms / (args.steps * nsmall)             # average duration of one full-patch pass, live
    achieved = bytes_local / (kernel_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": bytes_local,
                "kernel": {wrf.KERNEL_AUTO: "amt_pipe_kernel", wrf.KERNEL_PIPE: "amt_pipe_kernel", wrf.KERNEL_TILE: "amt_tile_kernel", wrf.KERNEL_COLUMN: "amt_column_kernel"}[kernel],
                "avg_launch_ms": kernel_ms,
                "frac_of_nominal_8TBs": achieved / 8000.0}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tr = json.load(open(prof)).get(args.workload)
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_launch"]
                roofline["traffic_source"] = tr.get("source")
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": WORKLOADS[args.workload][6],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {WORKLOADS[args.workload][7]}",
                   "grid": f"{g.ide}x{g.jde}x{g.kde}", "small_steps_per_step": nsmall,
                   "points_per_small_step": n3_global, "flags": "specified=T periodic_x=F nested=F",
                   "decomposition": f"{px}x{py} (i x j) patches, halo {HALO}",
                   "kernel": args.kernel, "l2": l2_note, "launch": exec_mode},
        "roofline": roofline, "gpu_launches": launches,
    }
    if clocks:
        line["clocks"] = clocks

    # ---- end to end through the reference-facing C-ABI call with HOST buffers (rank-local patch) ----
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        fields = {k: v for k, v in host.items()}
        h2d = sum(fields[n].nbytes for n in ("ww_1", "u", "u_1", "v", "v_1", "t", "t_1", "ft", "mu", "mut", "muu",
                                             "muv", "mu_tend", "msfuy", "msfvx_inv", "msftx", "msfty",
                                             "dnw", "fnm", "fnp", "rdnw"))
        h2d += fields["ww"].nbytes // pg.shape3[1]             # ww: level 1 only
        d2h = 4 * (3 * n3_local + 4 * n2_local)
        wrf.lib().wrfb200_set_default_kernel(kernel)
        wrf.call_with_fields(fields, pg, *scalars, nsteps=nsmall)          # warm-up: allocates the cached mirrors
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            wrf.call_with_fields(fields, pg, *scalars, nsteps=nsmall)      # H2D + nsmall launches + D2H + sync
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        line["e2e"] = {"value": n3_global * nsmall * e2e_steps / dt, "unit": UNIT,
                       "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                       "ms_per_step": 1e3 * dt / e2e_steps, "steps": e2e_steps,
                       "api": "wrfb200_advance_mu_t_loop(host arrays, nsteps=%d): upload, %d launches, download, sync"
                              % (nsmall, nsmall)}
        wrf.lib().wrfb200_release_cache()

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    if world == 1 and not args.no_cpu:
        r = cpu_reference_run(g, scalars, dx, steps=10, warmup=2)     # ~1 s wall = 10-30 core-seconds of CPU work
        line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}

    # ---- the repo's own CUDA-C kernel, recompiled for sm_100a, kernel-only like the reference's timer ----
    if world == 1 and not args.no_ref_cuda:
        try:
            line["ref_cuda_kernel"] = time_reference_cuda_kernel(g, scalars, host, dev)
        except Exception as e:                                  # reported, never fatal
            line["ref_cuda_kernel"] = {"unavailable": str(e)[:200]}

    if rank == 0:
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # The captured graph holds NCCL kernels; tearing the communicator down under it can block forever
        # (seen on the B200 box).  Everything is measured and printed: leave together and skip the teardown.
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    patch.close()
    del pinned_keep


def time_reference_cuda_kernel(g, scalars, host, dev, reps=5):
    """The reference's advance_mu_t_kernel (unmodified TU, -arch=sm_100a -fmad=false) on the same inputs,
    launch geometry of advance_mu_t_no_async.cu:54-55, scratch arrays in global memory."""
    import torch
    import wrf_model_cuda_sample_b200 as wrf
    from oracle import loader
    if not loader.have_ref_cuda():
        raise RuntimeError("oracle/_ref/libref_cuda_kernel.so not built")
    d = {n: torch.from_numpy(host[n]).to(dev) for n in wrf.FIELDS}
    scratch = {"wdtn": torch.empty_like(d["u"]), "dvdxi": torch.empty_like(d["u"]), "dmdt": torch.empty_like(d["mu"])}
    ptrs = {n: d[n].data_ptr() for n in d}
    sptr = {n: scratch[n].data_ptr() for n in scratch}
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        loader.reference_cuda_kernel(ptrs, sptr, g, scalars, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        loader.reference_cuda_kernel(ptrs, sptr, g, scalars, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n3, _ = g.updated_points()
    return {"value": n3 / (ms * 1e-3), "unit": UNIT, "ms_per_small_step": ms,
            "what": "reference advance_mu_t_kernel.cu, unmodified, nvcc -O3 -arch=sm_100a -fmad=false, "
                    "<<<(idim/64+1, jdim), 64>>>, kernel-only"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=("ours", "reference"), default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="conus3")
    ap.add_argument("--kernel", choices=("auto", "pipe", "tile", "column"), default="auto")
    ap.add_argument("--pgrid", default="", help="process grid PXxPY (default: j-slabs 1xN)")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="N>1: do not capture the step in a CUDA graph")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    if os.environ.get("BENCH_HANG_DUMP"):                       # debugging aid: dump all stacks after N seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_HANG_DUMP"]), exit=True)
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
