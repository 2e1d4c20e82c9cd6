"""The workloads of BASELINE.json / SURVEY.md section 8(d), as concrete inputs
shared by the CPU (port) and GPU parity tests and by bench.py.

Each workload is a function (physics_material_index) -> dict(settings,
geometry, states) for n tracks. Initial states are deterministic functions of
(n, seed) drawn with numpy, so every implementation gets the same inputs."""
import numpy as np

from harness import (EVENT_LIMIT_DISTANCE, EVENT_LIMIT_ENERGY,
                     EVENT_LIMIT_GRAMMAGE, EVENT_LIMIT_TIME, MODE_BACKWARD,
                     MODE_CSDA, MODE_DISABLED, MODE_FORWARD, MODE_MIXED,
                     MODE_RANDOMISED, MODE_STRAGGLED, MODE_WEIGHTED, STEP_CHECK,
                     STEP_RAW, layered_geometry, make_settings, make_states,
                     uniform_geometry)


def log_uniform(n, lo, hi, seed):
    rng = np.random.default_rng(seed)
    return np.exp(rng.uniform(np.log(lo), np.log(hi), n))


def isotropic(n, seed):
    rng = np.random.default_rng(seed)
    c = rng.uniform(-1, 1, n)
    phi = rng.uniform(0, 2 * np.pi, n)
    s = np.sqrt(1 - c * c)
    return np.stack([s * np.cos(phi), s * np.sin(phi), c], axis=1)


def headline(material, n, seed=0):
    """Forward, straggled, mixed scattering, 1 km standard rock, 1 TeV."""
    return dict(
        settings=make_settings(MODE_STRAGGLED, MODE_MIXED,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=1e3),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, 1e3))


def straight(material, n, seed=0, scattering=MODE_DISABLED, spectrum=False):
    """configs[0], examples/pumas/straight.c:110-153: backward, straggled,
    1 km standard rock, elevation 90 deg, final energy 1 GeV (or log-uniform
    1 GeV .. 1 TeV), random charge."""
    rng = np.random.default_rng(seed)
    if spectrum:
        kf = log_uniform(n, 1., 1e3, seed + 1)
        wf = kf * np.log(1e3) * 2
    else:
        kf = 1.
        wf = 2.
    charge = np.where(rng.uniform(size=n) > 0.5, 1., -1.)
    return dict(
        settings=make_settings(MODE_STRAGGLED, scattering,
                               direction=MODE_BACKWARD,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=1e3),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, kf, charge=charge, weight=wf,
                           direction=(0., 0., -1.)))


def csda_water(material, n, seed=0):
    """configs[1]: forward CSDA, no scattering, 1 TeV muons through 1 km of
    water (closed form path, pumas.c:4777)."""
    return dict(
        settings=make_settings(MODE_CSDA, MODE_DISABLED,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=1e3),
        geometry=uniform_geometry(material("Water")),
        states=make_states(n, 1e3))


def csda_magnet(material, n, seed=0, direction=MODE_FORWARD):
    """Closed form CSDA in a uniform magnetised medium: transport_csda_deflect
    and csda_magnetic_transport (pumas.c:4968-5168), both the small rotation
    expansion and the Larmor polynomials, both charges."""
    geo = uniform_geometry(material("Air"), density=1.205)
    for k, b in enumerate((0., 2e-5, -4e-5)):
        geo.layer[0].magnet[k] = b
    rng = np.random.default_rng(seed)
    charge = np.where(rng.uniform(size=n) > 0.5, 1., -1.)
    if direction == MODE_FORWARD:
        energy = log_uniform(n, 0.05, 1e3, seed + 1)
        limit = dict(event=EVENT_LIMIT_DISTANCE,
                     limit_distance=3e3)
    else:
        energy = log_uniform(n, 0.05, 1e2, seed + 1)
        limit = dict(event=EVENT_LIMIT_DISTANCE | EVENT_LIMIT_ENERGY,
                     limit_distance=3e3, limit_energy=2e2)
    return dict(
        settings=make_settings(MODE_CSDA, MODE_DISABLED, direction=direction,
                               **limit),
        geometry=geo,
        states=make_states(n, energy, charge=charge,
                           direction=isotropic(n, seed + 2)))


def csda_spectrum(material, n, seed=0, direction=MODE_FORWARD, time=True):
    """Closed form CSDA over 10 decades of energy with every limit armed
    (`time=False`: all but the proper time limit, which the lean CSDA kernel
    hands to the general code)."""
    return dict(
        settings=make_settings(
            MODE_CSDA, MODE_DISABLED, direction=direction,
            event=(EVENT_LIMIT_DISTANCE | EVENT_LIMIT_GRAMMAGE |
                   EVENT_LIMIT_ENERGY | (EVENT_LIMIT_TIME if time else 0)),
            limit_distance=300., limit_grammage=5e5,
            limit_energy=(1e-2 if direction == MODE_FORWARD else 1e5),
            limit_time=250.),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, log_uniform(n, 1e-3, 1e7, seed),
                           direction=isotropic(n, seed + 1)))


def rock_3km(material, n, seed=0, energy=1e4):
    """configs[2]: forward, straggled, mixed scattering, 3 km standard rock."""
    return dict(
        settings=make_settings(MODE_STRAGGLED, MODE_MIXED,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=3e3),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, energy))


def mixed_loss(material, n, seed=0, scattering=MODE_DISABLED,
               direction=MODE_FORWARD):
    """Mixed (hybrid) energy loss, pumas.c:5632-5695."""
    e = 1e3 if direction == MODE_FORWARD else 10.
    return dict(
        settings=make_settings(MODE_MIXED, scattering, direction=direction,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=1e3),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, e, direction=(0., 0., 1.)))


def geometry_example(material, n, seed=0, phase="detailed", thickness=100.,
                     elevation=60.):
    """configs[3], examples/pumas/geometry.c: rock [0, t) + exponential air
    [t, 1000) with a geomagnetic field, backward, energy limit.
    phase "detailed": Kf = 1 GeV, straggled + mixed up to 100 GeV
    (geometry.c:268-277); phase "fast": Kf = 100 GeV, mixed loss, no scattering
    up to 1e6 GeV (geometry.c:278-286)."""
    geo = layered_geometry(
        [0., thickness, 1e3],
        [dict(material=material("StandardRock"), density=2.65e3),
         dict(material=material("Air"), density=1.205, scale=12e3, eps=5e-2,
              magnet=(0., 2e-5, -4e-5))],
        step_type=STEP_RAW)
    c = np.cos(np.radians(90. - elevation))
    s = np.sqrt(1. - c * c)
    rng = np.random.default_rng(seed)
    charge = np.where(rng.uniform(size=n) > 0.5, 1., -1.)
    if phase == "detailed":
        settings = make_settings(MODE_STRAGGLED, MODE_MIXED,
                                 direction=MODE_BACKWARD,
                                 event=EVENT_LIMIT_ENERGY, limit_energy=1e2)
        kf = 1.
    else:
        settings = make_settings(MODE_MIXED, MODE_DISABLED,
                                 direction=MODE_BACKWARD,
                                 event=EVENT_LIMIT_ENERGY, limit_energy=1e6)
        kf = 1e2
    return dict(settings=settings, geometry=geo,
                states=make_states(n, kf, charge=charge, weight=2.,
                                   direction=(-s, 0., -c)))


GRID_ZENITH, GRID_AZIMUTH = 6, 8


def geometry_grid(material, n, seed=0, thickness=100.):
    """configs[3] as BASELINE.json states it: the detailed phase of
    examples/pumas/geometry.c (rock + exponential air, geomagnetic field,
    backward, straggled + mixed up to 100 GeV) over a zenith x azimuth grid of
    GRID_ZENITH x GRID_AZIMUTH viewing directions: bin b holds the
    n // (GRID_ZENITH * GRID_AZIMUTH) consecutive tracks [b per_bin,
    (b + 1) per_bin) (pumas_b200_tally_segments sums their weights; ranks shard
    the grid by bin and by track id). The remainder goes to the last bin."""
    w = geometry_example(material, n, seed=seed, thickness=thickness)
    n_bins = GRID_ZENITH * GRID_AZIMUTH
    per_bin = max(1, n // n_bins)
    b = np.minimum(np.arange(n) // per_bin, n_bins - 1)
    zenith = np.radians(5. + 70. * ((b // GRID_AZIMUTH) + 0.5) / GRID_ZENITH)
    azimuth = 2. * np.pi * ((b % GRID_AZIMUTH) + 0.5) / GRID_AZIMUTH
    s, c = np.sin(zenith), np.cos(zenith)
    w["states"]["direction"][:, 0] = -s * np.cos(azimuth)
    w["states"]["direction"][:, 1] = -s * np.sin(azimuth)
    w["states"]["direction"][:, 2] = -c
    return w


def layers_check(material, n, seed=0):
    """Three layers with exact boundary location (PUMAS_STEP_CHECK bisection,
    pumas.c:6424-6485), forward, isotropic directions, starting mid-stack."""
    geo = layered_geometry(
        [-50., 0., 30., 80.],
        [dict(material=material("Water")),
         dict(material=material("StandardRock")),
         dict(material=material("Water"), density=1.03e3)],
        step_type=STEP_CHECK)
    return dict(
        settings=make_settings(MODE_STRAGGLED, MODE_MIXED),
        geometry=geo,
        states=make_states(n, log_uniform(n, 0.5, 50., seed),
                           direction=isotropic(n, seed + 1),
                           position=(0., 0., 10.)))


def decay_randomised(material, n, seed=0):
    """Randomised decay + time limit in the stepping engine."""
    return dict(
        settings=make_settings(MODE_STRAGGLED, MODE_MIXED,
                               decay=MODE_RANDOMISED,
                               event=EVENT_LIMIT_TIME, limit_time=2e3),
        geometry=uniform_geometry(material("Water")),
        states=make_states(n, log_uniform(n, 0.2, 20., seed)))


def no_loss(material, n, seed=0):
    """Energy loss disabled, scattering only (pumas.c:5611-5625)."""
    return dict(
        settings=make_settings(MODE_DISABLED, MODE_MIXED,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=50.),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, log_uniform(n, 0.1, 1e3, seed)))


def tau_ice(material, n, seed=0):
    """configs[4]: tau, backward, straggled, weighted decay, ice
    (dump ice.tau)."""
    return dict(
        settings=make_settings(MODE_STRAGGLED, MODE_MIXED,
                               direction=MODE_BACKWARD, decay=MODE_WEIGHTED,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=300.),
        geometry=uniform_geometry(material("Ice")),
        states=make_states(n, log_uniform(n, 1e2, 1e6, seed), weight=1.,
                           direction=(0., 0., -1.)))


def vertex_stop(material, n, seed=0, direction=MODE_FORWARD):
    """Stop at the first vertex: every VERTEX bit in the event mask
    (pumas.c:5377-5381, 5399-5401), wide energy range."""
    from harness import EVENT_VERTEX_COULOMB, EVENT_VERTEX_DEL
    return dict(
        settings=make_settings(MODE_STRAGGLED, MODE_MIXED, direction=direction,
                               event=(EVENT_VERTEX_DEL | EVENT_VERTEX_COULOMB |
                                      EVENT_LIMIT_DISTANCE),
                               limit_distance=2e3),
        geometry=uniform_geometry(material("Water")),
        states=make_states(n, log_uniform(n, 1e-2, 1e5, seed),
                           direction=isotropic(n, seed + 1)))


def limits_all(material, n, seed=0, direction=MODE_FORWARD,
               energy_loss=MODE_STRAGGLED):
    """Every external limit armed in the stepping engine, energies from just
    above the first grid node up to 1 PeV, both directions."""
    k = log_uniform(n, 2e-3, 1e6, seed)
    return dict(
        settings=make_settings(
            energy_loss, MODE_MIXED, direction=direction,
            event=(EVENT_LIMIT_DISTANCE | EVENT_LIMIT_GRAMMAGE |
                   EVENT_LIMIT_ENERGY | EVENT_LIMIT_TIME),
            limit_distance=80., limit_grammage=3e5,
            limit_energy=(5e-2 if direction == MODE_FORWARD else 1e12),
            limit_time=60.),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, k[:n], direction=isotropic(n, seed + 1)))


def above_grid(material, n, seed=0, energy_loss=MODE_STRAGGLED):
    """Energies beyond the last grid node (1e9 GeV): the analytic
    extrapolations of pumas.c:3651-3660, 3833-3838, 3901, no scattering."""
    return dict(
        settings=make_settings(energy_loss, MODE_DISABLED,
                               event=EVENT_LIMIT_DISTANCE, limit_distance=10.),
        geometry=uniform_geometry(material("StandardRock")),
        states=make_states(n, log_uniform(n, 5e8, 1e11, seed)))


def dcs_models(material, n, seed=0, direction=MODE_FORWARD,
               medium="StandardRock"):
    """Physics built with non-default DCS models (pb_dcs_models.h): hard
    losses over five decades of energy, where the tabulated DCS falls back to
    the model (pumas.c:11279-11346) and the photonuclear polar angle is drawn
    from the model's doubly differential cross-section (pumas.c:11920)."""
    k = log_uniform(n, 1., 1e5, seed)
    limit = dict(limit_distance=300.)
    event = EVENT_LIMIT_DISTANCE
    if direction == MODE_BACKWARD:
        limit["limit_energy"] = 1e7
        event |= EVENT_LIMIT_ENERGY
    return dict(
        settings=make_settings(MODE_STRAGGLED, MODE_MIXED, direction=direction,
                               event=event, **limit),
        geometry=uniform_geometry(material(medium)),
        states=make_states(n, k, direction=isotropic(n, seed + 1)))


def with_density(material, n, seed=0, base=None, density=2.0e3, **kwargs):
    """A workload of a uniform medium with the density set by a (constant)
    locals callback, as examples/pumas/straight.c:61-74 does: the fast paths of
    the uniform medium must apply, with that density."""
    w = base(material, n, seed=seed, **kwargs)
    g = w["geometry"]
    w["geometry"] = uniform_geometry(g.layer[0].material, density=density)
    return w


# name -> (dump, builder, kwargs)
WORKLOADS = {
    "headline": ("materials.muon", headline, {}),
    "headline_density": ("materials.muon", with_density,
                         dict(base=headline, density=3.1e3)),
    "straight_density": ("materials.muon", with_density,
                         dict(base=straight, density=2.0e3)),
    "csda_density": ("materials.muon", with_density,
                     dict(base=csda_water, density=0.92e3)),
    "straight": ("materials.muon", straight, {}),
    "straight_mixed": ("materials.muon", straight,
                       dict(scattering=MODE_MIXED, spectrum=True)),
    "csda_water": ("materials.muon", csda_water, {}),
    "csda_spectrum_fwd": ("materials.muon", csda_spectrum, {}),
    "csda_spectrum_bwd": ("materials.muon", csda_spectrum,
                          dict(direction=MODE_BACKWARD)),
    "csda_limits_fwd": ("materials.muon", csda_spectrum, dict(time=False)),
    "csda_limits_bwd": ("materials.muon", csda_spectrum,
                        dict(direction=MODE_BACKWARD, time=False)),
    "csda_magnet_fwd": ("materials.muon", csda_magnet, {}),
    "csda_magnet_bwd": ("materials.muon", csda_magnet,
                        dict(direction=MODE_BACKWARD)),
    "models_kkb_fwd": ("rock.muon.kkb", dcs_models, {}),
    "models_kkb_bwd": ("rock.muon.kkb", dcs_models,
                       dict(direction=MODE_BACKWARD, medium="Water")),
    "models_abm_fwd": ("rock.muon.abm", dcs_models, dict(medium="Water")),
    "models_abm_bwd": ("rock.muon.abm", dcs_models,
                       dict(direction=MODE_BACKWARD)),
    "rock_3km": ("materials.muon", rock_3km, {}),
    "mixed_loss": ("materials.muon", mixed_loss, {}),
    "mixed_loss_scattering": ("materials.muon", mixed_loss,
                              dict(scattering=MODE_MIXED)),
    "mixed_loss_bwd": ("materials.muon", mixed_loss,
                       dict(direction=MODE_BACKWARD, scattering=MODE_MIXED)),
    "geometry_detailed": ("materials.muon", geometry_example, {}),
    "geometry_fast": ("materials.muon", geometry_example, dict(phase="fast")),
    "geometry_grid": ("materials.muon", geometry_grid, {}),
    "layers_check": ("materials.muon", layers_check, {}),
    "decay_randomised": ("materials.muon", decay_randomised, {}),
    "no_loss": ("materials.muon", no_loss, {}),
    "tau_ice": ("ice.tau", tau_ice, {}),
    "vertex_stop": ("materials.muon", vertex_stop, {}),
    "vertex_stop_bwd": ("materials.muon", vertex_stop,
                        dict(direction=MODE_BACKWARD)),
    "above_grid": ("materials.muon", above_grid, {}),
    "above_grid_mixed": ("materials.muon", above_grid,
                         dict(energy_loss=MODE_MIXED)),
    "limits_all": ("materials.muon", limits_all, {}),
    "limits_all_bwd": ("materials.muon", limits_all,
                       dict(direction=MODE_BACKWARD)),
    "limits_all_csda": ("materials.muon", limits_all,
                        dict(energy_loss=MODE_CSDA)),
    "limits_all_mixed_bwd": ("materials.muon", limits_all,
                             dict(energy_loss=MODE_MIXED,
                                  direction=MODE_BACKWARD)),
}
