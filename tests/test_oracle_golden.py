"""Pins the oracle (the reference compiled from its own sources, oracle/_ref)
against the known-answer values of the reference's own test suite
(tests/test-pumas.c:908-986), evaluated on the committed physics dump
(examples/data/materials.xml, muon).

One constant of that suite is not met by the reference as built here: the CSDA
stopping power of standard rock at 1 GeV is asserted to be 1.823E-04 +- 1E-06
(test-pumas.c:911) whereas src/pumas.c v1.2.3 tabulates 1.813E-04 (its own
generated standard_rock.txt has 1.813 MeV cm2/g at 1 GeV). The value pinned
below is the one the reference computes."""
import pytest

from harness import MODE_CSDA, MODE_MIXED

ROCK_DENSITY = 2.65e3


def check(rc_value, expected=None, tol=None, gt=None, lt=None):
    rc, value = rc_value
    assert rc == 0
    if expected is not None:
        assert abs(value - expected) <= tol, value
    if gt is not None:
        assert value > gt, value
    if lt is not None:
        assert value < lt, value
    return value


def test_property_known_answers(oracle):
    p = oracle.property
    check(p("cross_section", 0, 0.), expected=0., tol=0.)
    v = check(p("cross_section", 0, 1e3), gt=0.)
    assert 1. / (v * ROCK_DENSITY) < 1e3
    v = check(p("cross_section", 0, 1e6), gt=0.)
    assert 1. / (v * ROCK_DENSITY) < 1e3
    check(p("stopping_power", MODE_CSDA, 0, 1.), expected=1.813e-4, tol=1e-6)
    check(p("stopping_power", MODE_CSDA, 0, 1e3), expected=6.610e-4, tol=2e-7)
    check(p("stopping_power", MODE_MIXED, 0, 1e3), lt=6.610e-4)
    check(p("energy_straggling", 0, 1e3), expected=3.743e-4, tol=1e-7)
    check(p("range", MODE_CSDA, 0, 1.), expected=5.518e3, tol=1.)
    check(p("range", MODE_CSDA, 0, 1e3), expected=2.449e6, tol=1e3)
    check(p("range", MODE_MIXED, 0, 1e3), gt=2.449e6)
    check(p("kinetic_energy", MODE_CSDA, 0, 5.518e3), expected=1., tol=1e-3)
    check(p("kinetic_energy", MODE_CSDA, 0, 2.449e6), expected=1e3, tol=1.)
    check(p("kinetic_energy", MODE_MIXED, 0, 2.449e6), lt=1e3)
    check(p("magnetic_rotation", 0, 1.), gt=5.475e3 / 3.3374)
    check(p("proper_time", MODE_CSDA, 0, 1.), lt=5.475e3)
    check(p("proper_time", MODE_CSDA, 0, 1e-2), gt=8.498)
    check(p("elastic_path", 0, 1.), expected=275.89, tol=1e-2)
    check(p("elastic_cutoff_angle", 0, 1.), expected=0.00466, tol=1e-5)
    check(p("transport_path", MODE_MIXED, 0, 1.), expected=2.269e6, tol=1e3)


def test_materials(oracle):
    assert oracle.material("StandardRock") == 0
    assert oracle.material("Water") == 1
    assert oracle.material("Air") == 2
    assert oracle.material("WetRock") == 3
    with pytest.raises(KeyError):
        oracle.material("Unobtainium")
