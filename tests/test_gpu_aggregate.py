"""GPU parity, aggregate level (BASELINE.json north_star, criterion 2): fluxes,
transmission probabilities and final-state spectra of the CUDA path agree with
the reference within statistical error, on INDEPENDENT samples (different
Philox seeds on the two sides), so that a bias that per-track parity on shared
random numbers could hide (a correlated error) would show up.

Also: the device tally kernel (pumas_b200_tally, what the ranks of a multi-GPU
run all-reduce) against numpy on the same transported batch.
"""
import math
import os

import numpy as np
import pytest

import harness
import pumas_b200 as pb
from workloads import WORKLOADS

pytestmark = pytest.mark.gpu

# two-sided p-value below which a comparison fails: 1e-4 (3.9 sigma)
Z_MAX = 3.9
P_MIN = 1e-4


def _chi2_sf(x, k):
    """Survival function of chi2 with k degrees of freedom (regularised upper
    incomplete gamma, series / continued fraction; no scipy dependency)."""
    a, xx = 0.5 * k, 0.5 * x
    if xx <= 0:
        return 1.
    if xx < a + 1.:
        term = total = 1. / a
        n = a
        for _ in range(10000):
            n += 1.
            term *= xx / n
            total += term
            if abs(term) < abs(total) * 1e-15:
                break
        return max(0., 1. - total * math.exp(-xx + a * math.log(xx) -
                                             math.lgamma(a)))
    b = xx + 1. - a
    c = 1. / 1e-300
    d = 1. / b
    h = d
    for i in range(1, 10000):
        an = -i * (i - a)
        b += 2.
        d = an * d + b
        d = 1. / (d if abs(d) > 1e-300 else 1e-300)
        c = b + an / c
        c = c if abs(c) > 1e-300 else 1e-300
        delta = d * c
        h *= delta
        if abs(delta - 1.) < 1e-15:
            break
    return math.exp(-xx + a * math.log(xx) - math.lgamma(a)) * h


def _both(name, n, seed_ref, seed_gpu):
    if not harness.have_oracle():
        pytest.skip("oracle/_ref not built")
    assert pb.device_available(), "no B200 visible: the GPU tests need one"
    dump, builder, kwargs = WORKLOADS[name]
    oracle = harness.Oracle(dump)
    physics = pb.Physics(harness.dump_path(dump))
    w = builder(oracle.material, n, **kwargs)
    so, eo, _, _, rc, _ = oracle.transport(
        w["settings"], w["geometry"], w["states"], seed=seed_ref,
        threads=os.cpu_count() or 1)
    assert rc == 0
    ctx = pb.Context(physics, seed=seed_gpu)
    ctx.configure(w["settings"])
    ctx.geometry(w["geometry"])
    sg = w["states"].copy()
    eg, _ = ctx.transport(sg)
    return w, (so, eo), (sg, eg)


def _compare_spectra(a, b, bins, what):
    ha, _ = np.histogram(a, bins=bins)
    hb, _ = np.histogram(b, bins=bins)
    keep = (ha + hb) >= 20
    ha, hb = ha[keep].astype(float), hb[keep].astype(float)
    na, nb = ha.sum(), hb.sum()
    # homogeneity test of two histograms
    chi2 = (((ha * math.sqrt(nb / na) - hb * math.sqrt(na / nb)) ** 2) /
            (ha + hb)).sum()
    ndf = len(ha) - 1
    p = _chi2_sf(chi2, ndf)
    print("%s: chi2/ndf = %.1f/%d, p = %.3g" % (what, chi2, ndf, p))
    assert p > P_MIN, (what, chi2, ndf, p)


def test_transmission_and_spectrum_headline():
    """Survival probability through 1 km of rock and the spectrum of the
    survivors, forward straggled + mixed scattering at 1 TeV."""
    n = 20000
    w, (so, eo), (sg, eg) = _both("headline", n, seed_ref=101, seed_gpu=202)
    hit_o = (eo == pb.EVENT_LIMIT_DISTANCE)
    hit_g = (eg == pb.EVENT_LIMIT_DISTANCE)
    po, pg = hit_o.mean(), hit_g.mean()
    pool = 0.5 * (po + pg)
    z = (pg - po) / math.sqrt(2. * pool * (1. - pool) / n)
    print("survival: reference %.4f, GPU %.4f, z = %.2f" % (po, pg, z))
    assert abs(z) < Z_MAX
    bins = np.linspace(0., 1000., 26)
    _compare_spectra(so["energy"][hit_o], sg["energy"][hit_g], bins,
                     "final kinetic energy")
    # lateral displacement (multiple scattering)
    ro = np.hypot(so["position"][hit_o, 0], so["position"][hit_o, 1])
    rg = np.hypot(sg["position"][hit_g, 0], sg["position"][hit_g, 1])
    _compare_spectra(ro, rg, np.linspace(0., 6., 25), "lateral displacement")
    # range of the stopped tracks
    _compare_spectra(so["distance"][~hit_o], sg["distance"][~hit_g],
                     np.linspace(0., 1000., 21), "range of stopped tracks")


def test_backward_flux_estimate_straight():
    """examples/pumas/straight.c: the backward Monte Carlo estimate
    sum(w f(K)) of the flux through 1 km of rock, with a power law primary
    spectrum f (the estimate only needs to be the same functional on both
    sides), and the spectrum of the initial energies."""
    n = 40000
    w, (so, eo), (sg, eg) = _both("straight", n, seed_ref=303, seed_gpu=404)

    def estimate(s, e):
        ok = (e == pb.EVENT_LIMIT_DISTANCE)
        x = np.where(ok, s["weight"] * (s["energy"] + 0.10566) ** -2.7, 0.)
        return x.mean(), x.std(ddof=1) / math.sqrt(len(x))

    mo, so_ = estimate(so, eo)
    mg, sg_ = estimate(sg, eg)
    z = (mg - mo) / math.hypot(so_, sg_)
    print("flux estimate: reference %.4e +- %.1e, GPU %.4e +- %.1e, z = %.2f"
          % (mo, so_, mg, sg_, z))
    assert abs(z) < Z_MAX
    _compare_spectra(np.log10(so["energy"]), np.log10(sg["energy"]),
                     np.linspace(2., 4., 21), "log10 initial energy")


def test_tally_kernel_matches_numpy():
    """pumas_b200_tally on a transported batch: sum w, sum w^2, counts and the
    log-spaced energy histogram of the selected tracks."""
    import torch
    if not harness.have_oracle():
        pytest.skip("oracle/_ref not built")
    dump, builder, kwargs = WORKLOADS["straight"]
    physics = pb.Physics(harness.dump_path(dump))
    n = 5000
    w = builder(physics.material_index, n, **kwargs)
    ctx = pb.Context(physics, seed=5)
    ctx.configure(w["settings"])
    ctx.geometry(w["geometry"])
    states = w["states"].copy()
    events, _ = ctx.transport(states)

    n_bins, kmin, kmax = 24, 1e2, 1e4
    dev = torch.from_numpy(states.view(np.uint8).reshape(n, 104)).cuda()
    ev = torch.from_numpy(events.astype(np.int32)).cuda()
    out = torch.zeros(4 + n_bins, dtype=torch.float64, device="cuda")
    scratch = torch.zeros(pb.tally_scratch_bytes(n_bins), dtype=torch.uint8,
                          device="cuda")
    pb.tally(dev.data_ptr(), ev.data_ptr(), n, pb.EVENT_LIMIT_DISTANCE, kmin,
             kmax, n_bins, out.data_ptr(), scratch.data_ptr(),
             torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = out.cpu().numpy()

    sel = (events & pb.EVENT_LIMIT_DISTANCE) != 0
    wgt = states["weight"][sel]
    assert got[2] == n and got[3] == sel.sum()
    assert np.isclose(got[0], wgt.sum(), rtol=1e-12)
    assert np.isclose(got[1], (wgt ** 2).sum(), rtol=1e-12)
    k = states["energy"][sel]
    inside = (k >= kmin) & (k < kmax)
    idx = np.floor(np.log(k[inside] / kmin) / np.log(kmax / kmin) *
                   n_bins).astype(int)
    ref = np.bincount(np.clip(idx, 0, n_bins - 1), minlength=n_bins)
    assert np.abs(got[4:] - ref).sum() <= 2  # bin edges: libm rounding
    assert got[4:].sum() == inside.sum()
