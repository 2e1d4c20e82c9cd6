"""GPU parity report: every workload of tests/workloads.py through the C ABI
against the reference (oracle/_ref) on the same inputs and Philox streams.

Writes one JSON object: per workload n, match (within 1e-9 and in sync),
bitexact, diverged (different number of variates or end event), off_in_sync
(in sync but above 1e-9), max_rel over the in-sync tracks, steps/track and the
kernel time. Usage (on the GPU box):
    PUMAS_B200_LIBRARY=... python profiles/tools/parity_report.py OUT.json [scale]
`scale` multiplies the number of tracks of the parity tests (default 1)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

import harness  # noqa: E402
import pumas_b200 as pb  # noqa: E402
from workloads import WORKLOADS  # noqa: E402

N_TRACKS = {
    "headline": 100000, "rock_3km": 20000, "straight": 100000,
    "headline_density": 20000, "straight_density": 50000,
    "csda_density": 100000,
    "straight_mixed": 50000, "csda_water": 200000,
    "csda_spectrum_fwd": 100000, "csda_spectrum_bwd": 100000,
    "csda_limits_fwd": 100000, "csda_limits_bwd": 100000,
    "mixed_loss": 100000, "mixed_loss_scattering": 20000,
    "mixed_loss_bwd": 40000, "csda_magnet_fwd": 100000,
    "csda_magnet_bwd": 100000, "models_kkb_fwd": 20000,
    "models_kkb_bwd": 20000, "models_abm_fwd": 20000, "models_abm_bwd": 20000,
    "geometry_detailed": 40000, "geometry_fast": 100000,
    "geometry_grid": 48000,
    "layers_check": 20000, "decay_randomised": 20000, "no_loss": 10000,
    "tau_ice": 100000, "above_grid": 20000, "above_grid_mixed": 100000,
    "vertex_stop": 100000, "vertex_stop_bwd": 100000, "limits_all": 40000,
    "limits_all_bwd": 40000, "limits_all_csda": 40000,
    "limits_all_mixed_bwd": 40000,
}


def main():
    out_path = sys.argv[1]
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.
    only = sys.argv[3].split(",") if len(sys.argv) > 3 else None
    threads = os.cpu_count() or 1
    cache = {}
    report = {"library": pb.LIBRARY, "rtol": 1e-9, "workloads": {}}
    for name in sorted(WORKLOADS):
        if only and name not in only:
            continue
        dump, builder, kwargs = WORKLOADS[name]
        if dump not in cache:
            cache[dump] = (harness.Oracle(dump),
                           pb.Physics(harness.dump_path(dump)))
        oracle, physics = cache[dump]
        n = max(256, int(N_TRACKS.get(name, 20000) * scale))
        w = builder(oracle.material, n, **kwargs)
        t0 = time.time()
        so, eo, mo, ro, rc, _ = oracle.transport(
            w["settings"], w["geometry"], w["states"], seed=7, offset=1000,
            threads=threads)
        t_cpu = time.time() - t0
        ctx = pb.Context(physics, seed=7)
        ctx.configure(w["settings"])
        ctx.geometry(w["geometry"])
        ctx.track_offset(1000)
        sg = w["states"].copy()
        eg, mg = ctx.transport(sg)
        rg, steps = ctx.last_counters(n)
        ms, launches = ctx.last_kernel_time()
        c = harness.compare_states(so, sg, eo, eg, ro, rg, rtol=1e-9)
        sync = ~c["diverged"]
        off_in_sync = int(((c["worst"] > 1e-9) & sync).sum())
        media_ok = bool((mo[sync] == mg[sync]).all())
        # aggregate check on the diverged tracks: they are still samples of
        # the same distribution
        row = {
            "n": n, "match": c["n_match"], "bitexact": c["n_bitexact"],
            "diverged": c["n_diverged"], "off_in_sync": off_in_sync,
            "max_rel_in_sync": c["max_rel"], "media_equal": media_ok,
            "steps_per_track": float(steps.mean()),
            "kernel_ms": ms, "tracks_per_s_kernel": n / (ms * 1e-3) if ms else None,
            "cpu_reference_s": t_cpu, "cpu_threads": threads,
            "mean_energy_ref": float(so["energy"].mean()),
            "mean_energy_gpu": float(sg["energy"].mean()),
        }
        report["workloads"][name] = row
        print("%-22s n=%-7d match=%-7d bitexact=%-7d diverged=%-4d "
              "off_in_sync=%-4d max_rel=%.3g  %.1f ms" % (
                  name, n, c["n_match"], c["n_bitexact"], c["n_diverged"],
                  off_in_sync, c["max_rel"], ms), flush=True)
    with open(out_path, "w") as f:
        json.dump(report, f, indent=1)


if __name__ == "__main__":
    main()
