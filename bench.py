#!/usr/bin/env python3
"""Benchmark of the transport hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload ("headline"): forward transport of 1 TeV muons
through 1 km of standard rock, straggled energy loss + mixed scattering,
distance limit. A "step" is one pass of the batched transport entry point over
one batch of --tracks synthetic tracks per GPU. Tracks shard across ranks by
contiguous track id ranges (weak scaling: per GPU work is fixed); the tallies
(sum w, sum w^2, counts per energy bin) are accumulated on the device across
the steps and all-reduced ONCE at the end of the timed region, by the library
itself (pumas_b200_comm_allreduce, NCCL), as in BASELINE.json's north_star.
Nothing is waited for between the steps of the timed region
(pumas_b200_context_transport_async; one pumas_b200_context_synchronize).

One JSON line is printed by rank 0:
  value     whole-job tracks/s, inputs resident in HBM when the timed region
            starts (device pointers through the C ABI), CUDA events, max over
            ranks, barrier + synchronize on both sides;
  e2e       the same metric through the same C-ABI call with HOST (pinned)
            buffers: H2D of the states and D2H of states + events inside the
            timed region;
  roofline  FP64-pipe roofline of the transport kernel: algorithmic flop =
            600 flop per Monte Carlo step x the steps the launch
            executed (counted by the kernel), over the kernel's CUDA-event
            duration, against the DFMA peak measured live on this GPU;
  cpu_baseline  the reference itself (oracle/_ref, the unmodified src/pumas.c)
            on all host cores, one context per thread, on a bounded sample of
            the same workload (rank 0, N = 1 only).
`--impl reference` times that CPU reference alone with the same JSON schema.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "transported muon tracks/sec (straggled + mixed scattering, 1 km rock)"
UNIT = "tracks/s"
WORKLOAD = ("headline: forward, straggled energy loss, mixed scattering, "
            "StandardRock uniform, 1 TeV muons, distance limit 1000 m")
FLOP_PER_STEP = 600.
TALLY_BINS = 60


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tracks", type=int, default=1 << 23,
                    help="tracks per step and per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="tracks of the CPU baseline sample (0: auto)")
    ap.add_argument("--workload", default="headline")
    ap.add_argument("--no-extra", action="store_true",
                    help="skip the other BASELINE configurations")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,"
              "clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = threading.Event()

    def run(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", "-i", str(self.index),
                     "--query-gpu=" + self.FIELDS,
                     "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(
            ".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(
            ".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                 "sw_power_cap")
        reasons = [n for k, n in enumerate(names) if any(
            len(s) > 3 + k and s[3 + k].lower().startswith("active")
            for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


def reference_rate(workload, n, threads):
    """Tracks/s of the unmodified reference (oracle/_ref) on the host cores."""
    import harness
    from workloads import WORKLOADS
    dump, builder, kwargs = WORKLOADS[workload]
    oracle = harness.Oracle(dump)
    w = builder(oracle.material, n, **kwargs)
    out = oracle.transport(w["settings"], w["geometry"], w["states"], seed=1,
                           threads=threads)
    return n / out[5], float(out[3].mean())


def shared_config(n):
    """The `config` object, identical for the two arms (the reference arm times
    a bounded sample of the same workload: its size is under `details`)."""
    return {
        "workload": WORKLOAD,
        "rng": "Philox4x32-10 keyed by (seed, track id)",
        "sharding": "contiguous track id ranges per rank, no data-path "
                    "collective; tallies accumulated on the device and "
                    "all-reduced ONCE at the end of the run (NCCL, in-library)",
        "l2": "inputs larger than L2: %d MB of states per step and per GPU "
              "against 126 MB" % (n * 104 >> 20),
    }


def run_reference(args):
    """--impl reference: the CPU reference alone, same schema."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import harness
    if not harness.have_oracle():
        print(json.dumps({"impl": "reference", "unavailable":
                          "oracle/_ref/libpumas_oracle.so is not built"}))
        return
    threads = os.cpu_count() or 1
    # bounded sample: ~1.5 k tracks/s/core on this workload
    n = args.cpu_sample or max(2000, 1200 * threads)
    rates = []
    for _ in range(args.warmup):
        reference_rate(args.workload, max(200, n // 8), threads)
    t0 = time.time()
    for _ in range(args.steps):
        rates.append(reference_rate(args.workload, n, threads)[0])
    elapsed = time.time() - t0
    value = n * args.steps / sum(n / r for r in rates)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * elapsed / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": shared_config(args.tracks),
        "details": {"tracks_per_step": n,
                    "note": "CPU reference: unmodified src/pumas.c, one "
                            "pumas_context per thread, Philox callback; each "
                            "step is a bounded sample of the workload"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads,
                         "kind": "reference",
                         "sample": "%d tracks per step" % n},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# The other BASELINE.json configurations, measured after the timed region of
# the headline (one warm-up + three launches each, states resident in HBM; the
# MEDIAN launch counts: a backward batch now and then holds one track that runs
# away above the energy grid, where the rejection loop of the reference's hard
# elastic sampling (pumas.c:5843-5868) draws millions of variates, and the
# whole launch waits for it -- 1 track in 2e6 of `straight_mixed`):
# (workload of tests/workloads.py, tracks per launch, bound, algorithmic flop
# per Monte Carlo step or bytes per track).
EXTRA = [
    ("straight", 1 << 22, "fp64", 350.),          # configs[0], no scattering
    ("straight_mixed", 1 << 21, "fp64", 600.),    # configs[0], mixed
    ("csda_water", 1 << 23, "hbm", 212.),         # configs[1]
    ("rock_3km", 1 << 20, "fp64", 600.),          # configs[2]
    ("geometry_detailed", 1 << 19, "fp64", 600.),  # configs[3], one direction
    ("geometry_grid", 48 * 8192, "fp64", 600.),   # configs[3], 6 x 8 grid
    ("tau_ice", 1 << 21, "fp64", 600.),           # configs[4]
]
# pumas_b200_context_kernel_set (public knob): tracks of a few Monte Carlo steps
# (tau_ice: 3) spend their time entering and leaving the engine; the register
# resident kernel starts them in the stepping CTAs instead of the service CTAs
# of the shared memory resident kernel (B200, 4 M tracks: 7.2e6 against 5.6e6
# tracks/s; results within 1e-9 of the reference with either).
EXTRA_KERNEL = {"tau_ice": 3}


def measured_hbm_peak():
    """GB/s of MEASURED_PEAKS.json (driver written), else the fallback of
    B200_PROFILING.md."""
    try:
        return float(json.load(open(os.path.join(
            ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6650., "fallback"


def run_extra(pb, harness, WORKLOADS, torch, np, peak_tflops):
    out = {}
    hbm, hbm_kind = measured_hbm_peak()
    stream = torch.cuda.current_stream()
    cache = {}
    for name, n, bound, unit_cost in EXTRA:
        try:
            dump, builder, kwargs = WORKLOADS[name]
            if dump not in cache:
                cache[dump] = pb.Physics(harness.dump_path(dump))
            physics = cache[dump]
            w = builder(physics.material_index, n, **kwargs)
            ctx = pb.Context(physics, seed=2026)
            ctx.configure(w["settings"])
            ctx.geometry(w["geometry"])
            ctx.stream(stream.cuda_stream)
            if name in EXTRA_KERNEL:
                ctx.kernel(EXTRA_KERNEL[name])
            pristine = torch.from_numpy(
                w["states"].view(np.uint8).reshape(n, 104).copy()).cuda()
            work = torch.empty_like(pristine)
            events = torch.zeros(n, dtype=torch.int32, device="cuda")
            ms = []
            for k in range(4):
                work.copy_(pristine, non_blocking=True)
                ctx.track_offset(k * n)
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
                ctx.transport_async(n, work.data_ptr(), events.data_ptr())
                e1.record()
                ctx.synchronize()
                if k > 0:
                    ms.append(e0.elapsed_time(e1))
            _, nsteps = ctx.last_counters(n)
            steps = float(nsteps.mean())
            t = sorted(ms)[len(ms) // 2] * 1e-3
            row = {"tracks_per_s": n / t, "tracks_per_launch": n,
                   "launch_ms": [round(x, 3) for x in ms],
                   "steps_per_track": steps, "bound": bound,
                   "kernel": EXTRA_KERNEL.get(name, 0)}
            if bound == "hbm":
                row["achieved_GBps"] = unit_cost * n / t * 1e-9
                row["peak_GBps"] = hbm
                row["peak_kind"] = hbm_kind
                row["roofline_frac"] = row["achieved_GBps"] / hbm
            else:
                row["achieved_TFLOPs"] = unit_cost * steps * n / t * 1e-12
                row["roofline_frac"] = (row["achieved_TFLOPs"] / peak_tflops
                                        if peak_tflops else None)
            if name == "geometry_grid":
                # the tally of the grid: sum of the weights per viewing
                # direction (fixed order), what the ranks would all-reduce
                from workloads import GRID_AZIMUTH, GRID_ZENITH
                n_bins = GRID_ZENITH * GRID_AZIMUTH
                grid = torch.zeros(3 * n_bins, dtype=torch.float64,
                                   device="cuda")
                pb.tally_segments(work.data_ptr(), events.data_ptr(), n_bins,
                                  n // n_bins, pb.EVENT_LIMIT_ENERGY,
                                  grid.data_ptr(), stream.cuda_stream)
                g = grid.cpu().numpy().reshape(n_bins, 3)
                row["grid"] = {"zenith_bins": GRID_ZENITH,
                               "azimuth_bins": GRID_AZIMUTH,
                               "tracks_per_bin": n // n_bins,
                               "mean_weight_min": float((g[:, 0] / np.maximum(
                                   g[:, 2], 1.)).min()),
                               "mean_weight_max": float((g[:, 0] / np.maximum(
                                   g[:, 2], 1.)).max())}
            out[name] = row
            del ctx, pristine, work, events
        except Exception as exc:  # an extra must not lose the headline line
            out[name] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    return out


class QuietStdout:
    """stdout carries ONE JSON line: anything a library prints meanwhile (NCCL
    writes its version banner to stdout) goes to stderr instead."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    with QuietStdout():
        line = measure(args)
    if line is not None:
        print(json.dumps(line))


def measure(args):
    """Returns the JSON object on rank 0, None elsewhere."""

    import numpy as np
    import torch
    import torch.distributed as dist

    import harness
    import pumas_b200 as pb
    from workloads import WORKLOADS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available() or not pb.device_available():
        raise SystemExit("bench.py needs a B200: no usable CUDA device "
                         "(there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(
            "cuda", local_rank))

        def exchange(uid):
            box = [uid]
            dist.broadcast_object_list(box, src=0)
            return box[0]
        # the library's own communicator: the one collective of the data path
        comm = pb.Comm(rank, world, exchange)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.tracks
    dump, builder, kwargs = WORKLOADS[args.workload]
    physics = pb.Physics(harness.dump_path(dump))
    w = builder(physics.material_index, n, **kwargs)
    stream = torch.cuda.current_stream()
    ctx = pb.Context(physics, seed=2026)
    ctx.configure(w["settings"])
    ctx.geometry(w["geometry"])
    ctx.stream(stream.cuda_stream)

    # Inputs: pristine states in HBM (restored device-to-device before each
    # step), and in pinned host memory for the end-to-end leg.
    host = torch.from_numpy(
        w["states"].view(np.uint8).reshape(n, 104).copy()).pin_memory()
    pristine = host.cuda()
    work = torch.empty_like(pristine)
    events = torch.zeros(n, dtype=torch.int32, device="cuda")
    n_tally = 4 + TALLY_BINS
    tally = torch.zeros(n_tally, dtype=torch.float64, device="cuda")
    total = torch.zeros(n_tally, dtype=torch.float64, device="cuda")
    scratch = torch.zeros(pb.tally_scratch_bytes(TALLY_BINS), dtype=torch.uint8,
                          device="cuda")
    from pumas_b200.sharding import step_offset

    def run(first, count, stamps=None):
        """`count` passes of the hot path over this rank's batch, states
        resident in HBM, nothing waited for: every pass enqueues the restore
        of the inputs, the transport, the tally of the batch and its
        accumulation; then ONE all-reduce of the accumulated tallies."""
        total.zero_()
        for k in range(count):
            work.copy_(pristine, non_blocking=True)
            ctx.track_offset(step_offset(first + k, rank, world, n))
            if stamps is not None:
                stamps[k][0].record()
            ctx.transport_async(n, work.data_ptr(), events.data_ptr())
            if stamps is not None:
                stamps[k][1].record()
            pb.tally(work.data_ptr(), events.data_ptr(), n,
                     pb.EVENT_LIMIT_DISTANCE, 1e-3, 1e3, TALLY_BINS,
                     tally.data_ptr(), scratch.data_ptr(), stream.cuda_stream)
            pb.tally_accumulate(total.data_ptr(), tally.data_ptr(), n_tally,
                                stream.cuda_stream)
        if comm is not None:
            comm.allreduce(total.data_ptr(), n_tally, stream.cuda_stream)
        ctx.synchronize()

    run(0, args.warmup)
    barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler is not None:
        sampler.start()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    stamps = [(torch.cuda.Event(enable_timing=True),
               torch.cuda.Event(enable_timing=True))
              for _ in range(args.steps)]
    barrier()
    e0.record()
    run(args.warmup, args.steps, stamps)
    e1.record()
    barrier()
    if sampler is not None:
        sampler.stop_flag.set()
    elapsed_ms = e0.elapsed_time(e1)
    kernel_ms = sum(a.elapsed_time(b) for a, b in stamps)
    _, launches_per_step = ctx.last_kernel_time()
    # transport kernels + the two tally kernels + the accumulation
    launches = args.steps * (launches_per_step + 3)
    nrand, nsteps = ctx.last_counters(n)
    steps_per_track = float(nsteps.mean())
    totals = total.cpu().numpy()
    survival = float(totals[3] / totals[2]) if totals[2] else 0.

    # End to end: host (pinned) buffers through the reference-facing call.
    host_np = host.numpy().reshape(-1).view(pb.STATE_DTYPE)
    e2e_buf = torch.empty_like(host).pin_memory()
    e2e_np = e2e_buf.numpy().reshape(-1).view(pb.STATE_DTYPE)
    host_events = torch.zeros(n, dtype=torch.int32).pin_memory()

    def e2e_step(k):
        e2e_np[:] = host_np
        ctx.track_offset(step_offset(k, rank, world, n))
        t0 = time.perf_counter()
        ctx.transport(e2e_np, host_events.numpy(), media=False)
        return time.perf_counter() - t0

    e2e_step(0)
    barrier()
    e2e_steps = max(2, args.steps // 2)
    e2e_seconds = 0.
    for k in range(e2e_steps):
        e2e_seconds += e2e_step(100 + k)
    barrier()

    times = torch.tensor([elapsed_ms, e2e_seconds, kernel_ms],
                         dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    elapsed_ms, e2e_seconds, kernel_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        value = world * n * args.steps / (elapsed_ms * 1e-3)
        e2e_value = world * n * e2e_steps / e2e_seconds
        peak = pb.fp64_peak()
        flop_per_launch = FLOP_PER_STEP * steps_per_track * n
        achieved = flop_per_launch / (kernel_ms / args.steps * 1e-3) * 1e-12
        cpu = None
        extra = None
        if world == 1:
            if harness.have_oracle():
                threads = os.cpu_count() or 1
                sample = args.cpu_sample or max(2000, 1500 * threads)
                rate, _ = reference_rate(args.workload, sample, threads)
                cpu = {"value": rate, "unit": UNIT, "cores": threads,
                       "kind": "reference",
                       "sample": "%d tracks of the same workload, reference "
                                 "src/pumas.c, one context per thread" % sample}
            if not args.no_extra:
                extra = run_extra(pb, harness, WORKLOADS, torch, np, peak)
        config = shared_config(n)
        if extra is not None:
            config["extra"] = extra
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": config,
            "details": {
                "tracks_per_step_per_gpu": n,
                "steps_per_track": steps_per_track,
                "survival_fraction": survival,
                "tallies": "%d doubles per rank, %d collective(s) in the "
                           "timed region" % (n_tally, 1 if world > 1 else 0)},
            "clocks": sampler.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT,
                    "h2d_bytes_per_step": n * 104,
                    "d2h_bytes_per_step": n * 108},
            "gpu_launches": launches,
            "roofline": {
                "bound": "fp64", "achieved": achieved, "peak": peak,
                "unit": "TFLOP/s", "frac": achieved / peak if peak else None,
                "traffic": None,
                "note": "algorithmic flop = %.0f flop/step x %.1f steps/track "
                        "x %d tracks per launch over the CUDA-event duration "
                        "of the transport kernel; peak = DFMA peak measured "
                        "live (pumas_b200_fp64_peak; MEASURED_PEAKS.json has "
                        "no FP64 entry); HBM is not the bound (%.0f B/track "
                        "algorithmic); traffic: not measurable from inside "
                        "the run, see the ncu capture profiles/r2_hot_ncu.txt "
                        "(dram bytes per launch)" % (
                            FLOP_PER_STEP, steps_per_track, n, 212.)},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
    else:
        line = None
    if comm is not None:
        comm.close()
    if world > 1:
        dist.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
