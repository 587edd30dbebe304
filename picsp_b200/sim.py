"""Host-side mirror of the reference's hot-path interface over the C ABI.

Method names are the reference's own (/root/reference/src/main.cpp:202-231):
``scatterSpecies``, ``computeRho``, ``solvePotential``, ``spectralPotentialSolver``,
``computeEF``, ``pushSpecies``, ``rewindSpecies``, ``computeKE``, plus the loop order
of its ``main`` (``bootstrap`` = :453-472, ``step`` = :481-504).  Every method is one
call into libpicsp_b200.so; nothing is computed in Python.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from .lib import CParams, check, load_library

ION, ELECTRON = 0, 1
GRID_IDS = {"den_i": 0, "den_e": 1, "rho": 2, "phi": 3, "efx": 4, "efy": 5}
PHASES = ("deposit", "rho", "allreduce", "solve", "ef", "push", "sort", "step", "push_ions", "push_electrons")
FLAG_CLEAR_DENSITY, FLAG_NO_SORT, FLAG_NO_FUSE, FLAG_SOR_SINGLE_CTA, FLAG_SEPARATE_SORT, FLAG_NO_GRAPH, FLAG_WALLS, FLAG_NCCL_ONLY = 1, 2, 4, 8, 16, 32, 64, 128
FLAG_CUFFT_ONLY, FLAG_OWN_FFT = 256, 512

_dp = C.POINTER(C.c_double)


def _ptr(a):
    return a.ctypes.data_as(_dp)


@dataclass
class Params:
    """Normalised run parameters: the reference's globals after parse_ini_file (main.cpp:279-294)."""
    numxCells: int
    numyCells: int
    stepSize: float
    timeStep: float
    massI: float
    nParticlesI: int            # GLOBAL counts (spwt is defined on them, main.cpp:401-402)
    nParticlesE: int
    solverType: int = 1
    flags: int = 0
    device: int = 0
    capacity: tuple | None = None   # per-rank capacity; defaults to the global counts
    parts: int = 0                  # picsp_params::parts (0 = automatic: 1 unless the device memory asks for more)
    spwt: list = field(default_factory=list)

    def __post_init__(self):
        if not self.spwt:
            area = 1.0 * self.numxCells * self.numyCells * self.stepSize * self.stepSize
            self.spwt = [area / self.nParticlesI, area / self.nParticlesE]


class Simulation:
    def __init__(self, p: Params):
        self.L = load_library()
        self.p = p
        cp = CParams()
        cp.numxCells, cp.numyCells = p.numxCells, p.numyCells
        cp.stepSize, cp.timeStep = p.stepSize, p.timeStep
        cp.solverType, cp.flags = p.solverType, p.flags
        cp.charge[0], cp.charge[1] = 1.0, -1.0          # chargeE, -chargeE (main.cpp:407-408)
        cp.mass[0], cp.mass[1] = p.massI, 1.0
        cp.spwt[0], cp.spwt[1] = p.spwt
        cap = p.capacity or (p.nParticlesI, p.nParticlesE)
        cp.capacity[0], cp.capacity[1] = cap
        cp.device = p.device
        cp.parts = p.parts
        self.nix, self.niy = p.numxCells + 1, p.numyCells + 1
        self.ctx = C.c_void_p()
        check(self.L.picsp_create(C.byref(cp), C.byref(self.ctx)))

    def close(self):
        if getattr(self, "ctx", None):
            self.L.picsp_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- state exchange -----------------------------------------------------------------
    def set_species(self, s, x, y, vx, vy):
        arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, vx, vy)]
        check(self.L.picsp_species_upload(self.ctx, s, *(_ptr(a) for a in arrs), len(arrs[0])))

    def count(self, s):
        n = C.c_int64()
        check(self.L.picsp_species_count(self.ctx, s, C.byref(n)))
        return n.value

    def get_species(self, s):
        n = self.count(s)
        out = [np.empty(n) for _ in range(4)]
        check(self.L.picsp_species_download(self.ctx, s, *(_ptr(a) for a in out)))
        return tuple(out)

    def get_species_rows(self, s):
        rows = np.empty((self.count(s), 4))
        check(self.L.picsp_species_download_rows(self.ctx, s, _ptr(rows)))
        return rows

    def set_grid(self, name, a):
        a = np.ascontiguousarray(np.asarray(a, dtype=np.float64).reshape(-1))
        assert a.size == self.nix * self.niy
        check(self.L.picsp_grid_upload(self.ctx, GRID_IDS[name], _ptr(a)))

    def grid(self, name):
        out = np.empty(self.nix * self.niy)
        check(self.L.picsp_grid_download(self.ctx, GRID_IDS[name], _ptr(out)))
        return out

    def dump_begin(self, rows_i=None, rows_e=None, den_i=None, den_e=None, phi=None, ke=None):
        """picsp_dump_begin on caller-owned float64 buffers (numpy arrays or torch pinned tensors via .numpy());
        they must stay alive and untouched until dump_wait()."""
        self._dump_keep = (rows_i, rows_e, den_i, den_e, phi, ke)
        args = [(_ptr(a) if a is not None else None) for a in self._dump_keep]
        check(self.L.picsp_dump_begin(self.ctx, *args))

    def dump_wait(self):
        check(self.L.picsp_dump_wait(self.ctx))
        self._dump_keep = None

    def dump(self, root=True):
        """One whole diagnostics dump (what writeSpecies x2 + writePot + computeKE x2 produce, main.cpp:507-527)."""
        nn = self.nix * self.niy
        out = {"rows_i": np.empty((self.count(ION), 4)), "rows_e": np.empty((self.count(ELECTRON), 4)), "ke": np.empty(2)}
        if root:
            out.update(den_i=np.empty(nn), den_e=np.empty(nn), phi=np.empty(nn))
        self.dump_begin(out["rows_i"], out["rows_e"], out.get("den_i"), out.get("den_e"), out.get("phi"), out["ke"])
        self.dump_wait()
        return out

    def fill_synthetic(self, s, n, first_index=0, seed=0, vth=1.0, xdrift=0.0):
        check(self.L.picsp_species_fill_synthetic(self.ctx, s, n, first_index, seed, vth, xdrift))

    # -- the reference's function names ----------------------------------------------------
    def scatterSpecies(self, s): check(self.L.picsp_deposit(self.ctx, s))
    def computeRho(self): check(self.L.picsp_compute_rho(self.ctx))
    def spectralPotentialSolver(self): check(self.L.picsp_solve_spectral(self.ctx)); return True

    def solvePotential(self):
        sw, l2 = C.c_int64(), C.c_double()
        check(self.L.picsp_solve_sor(self.ctx, C.byref(sw), C.byref(l2)))
        self.last_sweeps, self.last_l2 = sw.value, l2.value
        return True

    def solve(self): check(self.L.picsp_solve(self.ctx))

    def solve_status(self):
        sw, l2 = C.c_int64(), C.c_double()
        check(self.L.picsp_solve_status(self.ctx, C.byref(sw), C.byref(l2)))
        return sw.value, l2.value
    def computeEF(self): check(self.L.picsp_compute_ef(self.ctx))
    def pushSpecies(self, s): check(self.L.picsp_push(self.ctx, s))
    def rewindSpecies(self, s): check(self.L.picsp_rewind(self.ctx, s))
    def bootstrap(self): check(self.L.picsp_bootstrap(self.ctx))
    def step(self, nsteps=1): check(self.L.picsp_step(self.ctx, nsteps))
    def sync(self): check(self.L.picsp_sync(self.ctx))

    def computeKE(self, s):
        ke = C.c_double()
        check(self.L.picsp_compute_ke(self.ctx, s, C.byref(ke)))
        return ke.value

    def delta_phi(self):
        m, p0 = C.c_double(), C.c_double()
        check(self.L.picsp_delta_phi(self.ctx, C.byref(m), C.byref(p0)))
        return m.value - p0.value

    def repush_count(self, s):
        n = C.c_int64()
        check(self.L.picsp_repush_count(self.ctx, s, C.byref(n)))
        return n.value

    def straggler_count(self, s):
        n = C.c_int64()
        check(self.L.picsp_straggler_count(self.ctx, s, C.byref(n)))
        return n.value

    def set_sort_period(self, s, period): check(self.L.picsp_set_sort_period(self.ctx, s, period))
    def set_cell_sort_period(self, s, period): check(self.L.picsp_set_cell_sort_period(self.ctx, s, period))
    def set_bank_order(self, s, mode): check(self.L.picsp_set_bank_order(self.ctx, s, mode))
    def set_deposit_aggregation(self, s, mode): check(self.L.picsp_set_deposit_aggregation(self.ctx, s, mode))

    # -- multi-GPU ------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = C.create_string_buffer(128)
        check(load_library().picsp_comm_unique_id(buf))
        return buf.raw

    def comm_attach(self, unique_id: bytes, rank: int, nranks: int):
        buf = C.create_string_buffer(unique_id, 128)
        check(self.L.picsp_comm_attach(self.ctx, buf, rank, nranks))

    def comm_peer_reduction(self):
        return bool(self.L.picsp_comm_peer_reduction(self.ctx))

    # -- instrumentation ------------------------------------------------------------------------
    def profile_enable(self, on=True): check(self.L.picsp_profile_enable(self.ctx, 1 if on else 0))
    def profile_reset(self): check(self.L.picsp_profile_reset(self.ctx))

    def profile(self):
        out = {}
        for i, name in enumerate(PHASES):
            ms, calls = C.c_double(), C.c_int64()
            check(self.L.picsp_profile_get(self.ctx, i, C.byref(ms), C.byref(calls)))
            out[name] = (ms.value, calls.value)
        return out

    def spectral_engine(self):
        """'own' (shared-memory DFT of fft_kernels.cuh) or 'cufft'."""
        n = C.c_int()
        check(self.L.picsp_spectral_engine(self.ctx, C.byref(n)))
        return "own" if n.value else "cufft"

    def parts(self):
        n = C.c_int()
        check(self.L.picsp_parts(self.ctx, C.byref(n)))
        return n.value

    def kernel_launches(self):
        n = C.c_int64()
        check(self.L.picsp_kernel_launches(self.ctx, C.byref(n)))
        return n.value


def shard_range(n_global: int, rank: int, nranks: int):
    """Particle index range [lo, hi) owned by a rank: loader order, contiguous blocks."""
    lo = (n_global * rank) // nranks
    hi = (n_global * (rank + 1)) // nranks
    return lo, hi
