"""CPU: live comparison of the C restatement with the UNMODIFIED reference translation unit
(oracle/_ref/libpicsp_ref.so).  Skipped where that library has not been built."""
import numpy as np
import pytest

from oracle.oracle import ELECTRON, ION, Oracle, Reference, have_reference, normalise
from tests.helpers import GRIDS

pytestmark = pytest.mark.skipif(not have_reference(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("solver,numx,n", [(2, 24, 3000), (1, 24, 3000), (1, 41, 2000), (2, 64, 5000)])
def test_loop_bitwise_vs_reference(solver, numx, n):
    nm = normalise()
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    r = Reference(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    try:
        r.seed(7); o.seed(7)
        for s in (ION, ELECTRON):
            r.init(s, 1); o.init(s, 1)
        # give ions electron-like speeds too so that both species cross cells and wrap
        x, y, vx, vy = o.get_species(ION)
        o.set_species(ION, x, y, vx * 30, vy * 30); r.set_species(ION, x, y, vx * 30, vy * 30)
        r.bootstrap(); o.bootstrap()
        for st in range(5):
            for g in GRIDS:
                assert np.array_equal(o.grid(g), r.grid(g)), (st, g)
            for s in (ION, ELECTRON):
                assert np.array_equal(np.stack(o.get_species(s)), np.stack(r.get_species(s))), (st, s)
                assert o.computeKE(s) == r.computeKE(s)
            r.step(1); o.step(1)
    finally:
        r.close()


@pytest.mark.parametrize("solver,numx,numy,n", [(1, 20, 36, 2500), (2, 36, 20, 2500), (1, 33, 18, 2000), (2, 17, 50, 2000)])
def test_rectangular_loop_bitwise_vs_reference(solver, numx, numy, n):
    """numxCells != numyCells (odd and even node counts in either direction), hot particles of both species so that
    wraps and re-pushes happen in x and in y: grids, phase space and KE stay bit-identical over 6 steps."""
    nm = normalise()
    o = Oracle(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    r = Reference(numx, numy, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    try:
        r.seed(11); o.seed(11)
        for s in (ION, ELECTRON):
            r.init(s, 1); o.init(s, 1)
        for s, f in ((ION, 40.0), (ELECTRON, 2.0)):
            x, y, vx, vy = o.get_species(s)
            o.set_species(s, x, y, vx * f, vy * f); r.set_species(s, x, y, vx * f, vy * f)
        r.bootstrap(); o.bootstrap()
        for st in range(6):
            r.step(1); o.step(1)
            for g in GRIDS:
                assert np.array_equal(o.grid(g), r.grid(g)), (st, g)
            for s in (ION, ELECTRON):
                assert np.array_equal(np.stack(o.get_species(s)), np.stack(r.get_species(s))), (st, s)
                assert o.computeKE(s) == r.computeKE(s)
    finally:
        r.close()
