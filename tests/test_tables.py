"""The dump loader (pb_dump.hpp) exposes the reference's tables bit for bit."""
import numpy as np
import pytest

import harness

TABLES = ["K", "K_dX", "X", "X_dK", "T", "dE", "Omega", "CS", "NI_in", "NI_el",
          "Mu0", "Lb", "Ms1", "Kt", "a_max", "b_max"]


@pytest.mark.parametrize("dump", ["materials.muon", "ice.tau"])
def test_tables_identical(dump):
    if not (harness.have_oracle() and harness.have_port()):
        pytest.skip("oracle/_ref or oracle/_port not built")
    oracle, port = harness.Oracle(dump), harness.Port(dump)
    for name in TABLES:
        a, b = oracle.table(name), port.table(name)
        assert a.shape == b.shape and a.size > 0
        assert np.array_equal(a, b), name


def test_grids_are_strictly_increasing():
    """The hinted search of pb_tables.h (locate_near) equals the reference's
    bisection only on strictly increasing tables: K and every X column."""
    if not harness.have_port():
        pytest.skip("oracle/_port not built")
    for dump in ("materials.muon", "ice.tau"):
        port = harness.Port(dump)
        k = port.table("K")
        assert (np.diff(k) > 0).all()
        x = port.table("X").reshape(-1, len(k))
        assert (np.diff(x, axis=1) > 0).all()
