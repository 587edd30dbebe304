"""oracle/oracle.py — TEST INFRASTRUCTURE ONLY.

ctypes front-ends for the two CPU checkers of the hot path:

* ``Oracle``     – oracle/libpicsp_oracle.so, the plain-C restatement (picsp_oracle.c)
* ``Reference``  – oracle/_ref/libpicsp_ref.so, the UNMODIFIED reference translation
                   unit (/root/reference/src/main.cpp) behind oracle/ref_harness.cpp

Both expose the reference's own function names (``scatterSpecies``, ``computeRho``,
``solvePotential``, ``spectralPotentialSolver``, ``computeEF``, ``pushSpecies``,
``rewindSpecies``, ``computeKE``; /root/reference/src/main.cpp:202-231) on numpy
state, so parity tests read like calls into the reference.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import this.
The product package (picsp_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "libpicsp_oracle.so")
REF_SO = os.path.join(HERE, "_ref", "libpicsp_ref.so")
REF_O0_SO = os.path.join(HERE, "_ref", "libpicsp_ref_O0.so")
REF_GPU_SO = os.path.join(HERE, "_ref", "libpicsp_ref_gpu.so")              # patched reference TU, per-function ABI calls
REF_GPU_FUSED_SO = os.path.join(HERE, "_ref", "libpicsp_ref_gpu_fused.so")  # same with picsp_step
REFERENCE_ROOT = os.environ.get("PICSP_REFERENCE_ROOT", "/root/reference")

ION, ELECTRON = 0, 1
_dp = C.POINTER(C.c_double)


def build(ref: bool | None = None) -> None:
    """Compile the C restatement and, when the reference tree is mounted, oracle/_ref."""
    subprocess.check_call(["make", "-s", "-C", HERE, "oracle"])
    if ref is None:
        ref = os.path.isfile(os.path.join(REFERENCE_ROOT, "src", "main.cpp"))
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref", f"REFERENCE={REFERENCE_ROOT}"])
        # the boundary proof (reference TU + INTEGRATION.md patch, linked against the product library)
        if os.path.isfile(os.path.join(os.path.dirname(HERE), "picsp_b200", "libpicsp_b200.so")):
            subprocess.check_call(["make", "-s", "-C", HERE, "refgpu", f"REFERENCE={REFERENCE_ROOT}"])


def have_reference() -> bool:
    return os.path.isfile(REF_SO)


def _ptr(a: np.ndarray):
    assert a.dtype == np.float64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_dp)


class _Domain(C.Structure):
    _fields_ = [("nix", C.c_int), ("niy", C.c_int), ("dx", C.c_double), ("dy", C.c_double),
                ("xl", C.c_double), ("yl", C.c_double), ("dt", C.c_double)]


def normalise(time_step=1e-10, step_size=1.2e-4, charge=1.602e-19, mass_e=9.109e-31,
              mass_i=1.673e-27, density=1e12, vth_e=0.9, vth_i=0.026, drift_e=0.2, drift_i=0.0):
    """The reference's unit normalisation (main.cpp:279-291) on its shipped physical values."""
    EPS_un, K, EV_TO_K = 8.85418782e-12, 1.38065e-23, 11604.52
    omega_pe = np.sqrt((charge * charge * density) / (mass_e * EPS_un))
    lambda_d = np.sqrt((EPS_un * K * vth_e * EV_TO_K) / (density * charge * charge))
    return dict(dt=float(time_step * omega_pe), dx=float(step_size / lambda_d),
                mass_i=mass_i / mass_e, vth_e=vth_e / vth_e, vth_i=vth_i / vth_e,
                drift_e=drift_e / vth_e, drift_i=drift_i / vth_e)


class _Base:
    """State layout shared by both checkers (numpy, float64)."""

    def __init__(self, numx, numy, dx, dt, mass_i, n_i, n_e, vth_i=0.0288888888888889, vth_e=1.0, solver=1):
        self.numx, self.numy = int(numx), int(numy)
        self.nix, self.niy = self.numx + 1, self.numy + 1
        self.dx, self.dt = float(dx), float(dt)
        self.n = [int(n_i), int(n_e)]
        self.mass = [float(mass_i), 1.0]
        self.charge = [1.0, -1.0]
        self.vth = [float(vth_i), float(vth_e)]
        self.solver = int(solver)
        # spwt exactly as main.cpp:401-402: (density*numxCells*numyCells*stepSize*stepSize)/N
        self.spwt = [(1.0 * self.numx * self.numy * self.dx * self.dx) / n for n in self.n]


class Oracle(_Base):
    """The plain-C restatement (oracle/picsp_oracle.c)."""

    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.isfile(ORACLE_SO):
                build(ref=False)
            L = C.CDLL(ORACLE_SO)
            D = C.POINTER(_Domain)
            L.oracle_domain_init.argtypes = [D, C.c_int, C.c_int, C.c_double, C.c_double]
            L.oracle_fold_periodic.argtypes = [D, _dp]
            L.oracle_scatter_only.argtypes = [D, _dp, _dp, _dp, C.c_long, C.c_double]
            L.oracle_deposit.argtypes = [D, _dp, _dp, _dp, C.c_long, C.c_double]
            L.oracle_compute_rho.argtypes = [D, _dp, _dp, _dp, C.c_double, C.c_double]
            L.oracle_sor.argtypes = [D, _dp, _dp, _dp]
            L.oracle_sor.restype = C.c_long
            L.oracle_spectral.argtypes = [D, _dp, _dp]
            L.oracle_compute_ef.argtypes = [D, _dp, _dp, _dp]
            L.oracle_push.argtypes = [D, _dp, _dp, _dp, _dp, _dp, _dp, C.c_long, C.c_double, C.c_double]
            L.oracle_push.restype = C.c_long
            L.oracle_rewind.argtypes = [D, _dp, _dp, _dp, _dp, _dp, _dp, C.c_long, C.c_double, C.c_double]
            L.oracle_compute_ke.argtypes = [_dp, _dp, C.c_long, C.c_double, C.c_double]
            L.oracle_compute_ke.restype = C.c_double
            L.oracle_delta_phi.argtypes = [D, _dp]
            L.oracle_delta_phi.restype = C.c_double
            L.oracle_mt_seed.argtypes = [C.c_void_p, C.c_uint32]
            L.oracle_rnd.argtypes = [C.c_void_p]
            L.oracle_rnd.restype = C.c_double
            L.oracle_load.argtypes = [D, C.c_int, C.c_void_p, _dp, C.c_long, C.c_double, C.c_double,
                                      C.c_double, _dp, _dp, _dp, _dp]
            L.oracle_set_fft_mode.argtypes = [C.c_int]
            cls._lib = L
        return cls._lib

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        L = self.lib()
        self.dom = _Domain()
        L.oracle_domain_init(C.byref(self.dom), self.numx, self.numy, self.dx, self.dt)
        nn = self.nix * self.niy
        self.guard = 4 * self.niy + 8
        self.den = [np.zeros(nn), np.zeros(nn)]
        self.rho = np.zeros(nn)
        self.phi = np.zeros(nn)
        self._efx = np.zeros(nn + 2 * self.guard)
        self._efy = np.zeros(nn + 2 * self.guard)
        self.efx = self._efx[self.guard:self.guard + nn]
        self.efy = self._efy[self.guard:self.guard + nn]
        self.x = [np.zeros(n) for n in self.n]
        self.y = [np.zeros(n) for n in self.n]
        self.vx = [np.zeros(n) for n in self.n]
        self.vy = [np.zeros(n) for n in self.n]
        self._mt = C.create_string_buffer(624 * 4 + 16)
        self._xcarry = np.zeros(1)
        self.seed(0)
        self.extra_pushes = 0

    # -- state exchange ------------------------------------------------------
    def set_species(self, s, x, y, vx, vy):
        self.x[s] = np.ascontiguousarray(x, dtype=np.float64).copy()
        self.y[s] = np.ascontiguousarray(y, dtype=np.float64).copy()
        self.vx[s] = np.ascontiguousarray(vx, dtype=np.float64).copy()
        self.vy[s] = np.ascontiguousarray(vy, dtype=np.float64).copy()
        self.n[s] = len(self.x[s])

    def get_species(self, s):
        return self.x[s].copy(), self.y[s].copy(), self.vx[s].copy(), self.vy[s].copy()

    def set_grid(self, name, a):
        self.grid(name)[...] = np.asarray(a, dtype=np.float64).reshape(-1)

    def grid(self, name):
        if name == "den_i": return self.den[0]
        if name == "den_e": return self.den[1]
        return getattr(self, name)

    # -- the reference's function names --------------------------------------
    def seed(self, seed):
        self.lib().oracle_mt_seed(self._mt, seed)
        self._xcarry[0] = 0.0

    def rnd(self):
        return self.lib().oracle_rnd(self._mt)

    def init(self, s, load_type, xdrift=0.0, ydrift=0.0):
        n = self.n[s]
        self.x[s], self.y[s], self.vx[s], self.vy[s] = (np.zeros(n) for _ in range(4))
        self.lib().oracle_load(C.byref(self.dom), load_type, self._mt, _ptr(self._xcarry), n, self.vth[s],
                               xdrift, ydrift, _ptr(self.x[s]), _ptr(self.y[s]), _ptr(self.vx[s]), _ptr(self.vy[s]))

    def scatterSpecies(self, s):
        self.lib().oracle_deposit(C.byref(self.dom), _ptr(self.den[s]), _ptr(self.x[s]), _ptr(self.y[s]),
                                  self.n[s], self.spwt[s])

    def computeRho(self):
        self.lib().oracle_compute_rho(C.byref(self.dom), _ptr(self.rho), _ptr(self.den[0]), _ptr(self.den[1]),
                                      self.charge[0], self.charge[1])

    def solvePotential(self):
        l2 = C.c_double(0)
        sweeps = self.lib().oracle_sor(C.byref(self.dom), _ptr(self.phi), _ptr(self.rho),
                                       C.cast(C.byref(l2), _dp))
        self.last_l2, self.last_sweeps = l2.value, sweeps
        return sweeps > 0

    def spectralPotentialSolver(self):
        self.lib().oracle_spectral(C.byref(self.dom), _ptr(self.phi), _ptr(self.rho))
        return True

    def solve(self):
        return self.spectralPotentialSolver() if self.solver == 1 else self.solvePotential()

    def _ef_ptrs(self):
        off = self.guard * 8
        return (C.cast(self._efx.ctypes.data + off, _dp), C.cast(self._efy.ctypes.data + off, _dp))

    def computeEF(self):
        ex, ey = self._ef_ptrs()
        self.lib().oracle_compute_ef(C.byref(self.dom), _ptr(self.phi), ex, ey)

    def pushSpecies(self, s):
        ex, ey = self._ef_ptrs()
        self.extra_pushes = self.lib().oracle_push(C.byref(self.dom), ex, ey, _ptr(self.x[s]), _ptr(self.y[s]),
                                                   _ptr(self.vx[s]), _ptr(self.vy[s]), self.n[s],
                                                   self.charge[s], self.mass[s])
        return self.extra_pushes

    def rewindSpecies(self, s):
        ex, ey = self._ef_ptrs()
        self.lib().oracle_rewind(C.byref(self.dom), ex, ey, _ptr(self.x[s]), _ptr(self.y[s]),
                                 _ptr(self.vx[s]), _ptr(self.vy[s]), self.n[s], self.charge[s], self.mass[s])

    def computeKE(self, s):
        return self.lib().oracle_compute_ke(_ptr(self.vx[s]), _ptr(self.vy[s]), self.n[s], self.spwt[s], self.mass[s])

    def delta_phi(self):
        return self.lib().oracle_delta_phi(C.byref(self.dom), _ptr(self.phi))

    # -- loop order, main.cpp:453-472 and :481-504 ----------------------------
    def bootstrap(self):
        self.scatterSpecies(ION); self.scatterSpecies(ELECTRON)
        self.computeRho(); self.solve(); self.computeEF()
        self.rewindSpecies(ION); self.rewindSpecies(ELECTRON)

    def step(self, nsteps=1):
        for _ in range(nsteps):
            self.scatterSpecies(ION); self.scatterSpecies(ELECTRON)
            self.computeRho(); self.solve(); self.computeEF()
            self.pushSpecies(ION); self.pushSpecies(ELECTRON)


class Reference(_Base):
    """The unmodified reference TU (global state: one live instance at a time)."""

    _libs = {}

    @classmethod
    def lib(cls, path=REF_SO):
        if path not in cls._libs:
            if not os.path.isfile(path):
                raise FileNotFoundError(f"{path} not built (run `make -C oracle ref` where /root/reference is mounted)")
            L = C.CDLL(path)
            L.picsp_ref_setup.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_int, C.c_int,
                                          C.c_double, C.c_double, C.c_int]
            L.picsp_ref_spwt.argtypes = [C.c_int]; L.picsp_ref_spwt.restype = C.c_double
            L.picsp_ref_grid_ptr.argtypes = [C.c_int]; L.picsp_ref_grid_ptr.restype = _dp
            L.picsp_ref_species_count.argtypes = [C.c_int]; L.picsp_ref_species_count.restype = C.c_long
            L.picsp_ref_species_set.argtypes = [C.c_int, _dp, _dp, _dp, _dp, C.c_long]
            L.picsp_ref_species_get.argtypes = [C.c_int, _dp, _dp, _dp, _dp]
            for f in ("scatterSpecies", "scatterSpeciesVel", "pushSpecies", "rewindSpecies"):
                getattr(L, "picsp_ref_" + f).argtypes = [C.c_int]
            L.picsp_ref_computeKE.argtypes = [C.c_int]; L.picsp_ref_computeKE.restype = C.c_double
            L.picsp_ref_step.argtypes = [C.c_int, C.c_int]
            L.picsp_ref_phase_seconds.argtypes = [_dp, C.c_int]
            L.picsp_ref_seed.argtypes = [C.c_uint]
            L.picsp_ref_rnd.restype = C.c_double
            L.picsp_ref_init.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double]
            L.picsp_ref_init_both.argtypes = [C.c_int, C.c_double, C.c_double]
            L.picsp_ref_parse_ini.argtypes = [C.c_char_p, _dp]
            L.picsp_ref_main.argtypes = [C.c_char_p]
            L.picsp_ref_h5_count.restype = C.c_long
            L.picsp_ref_h5_name.argtypes = [C.c_long]; L.picsp_ref_h5_name.restype = C.c_char_p
            L.picsp_ref_h5_meta.argtypes = [C.c_long, C.POINTER(C.c_longlong)]
            L.picsp_ref_h5_data.argtypes = [C.c_long]; L.picsp_ref_h5_data.restype = C.c_void_p
            L.picsp_ref_h5_group_count.restype = C.c_long
            L.picsp_ref_h5_group_name.argtypes = [C.c_long]; L.picsp_ref_h5_group_name.restype = C.c_char_p
            L.picsp_ref_set_fft_mode.argtypes = [C.c_int]
            cls._libs[path] = L
        return cls._libs[path]

    _GRID = {"den_i": 0, "den_e": 1, "rho": 2, "phi": 3, "efx": 4, "efy": 5}

    def __init__(self, *a, lib_path=REF_SO, **k):
        super().__init__(*a, **k)
        self.L = self.lib(lib_path)
        self.L.picsp_ref_setup(self.numx, self.numy, self.dx, self.dt, self.mass[0], self.n[0], self.n[1],
                               self.vth[0], self.vth[1], self.solver)
        assert self.L.picsp_ref_spwt(0) == self.spwt[0] and self.L.picsp_ref_spwt(1) == self.spwt[1]

    def grid(self, name):
        """A live numpy view of the reference's own array."""
        p = self.L.picsp_ref_grid_ptr(self._GRID[name])
        return np.ctypeslib.as_array(p, shape=(self.nix * self.niy,))

    def set_grid(self, name, a):
        self.grid(name)[...] = np.asarray(a, dtype=np.float64).reshape(-1)

    @property
    def den(self): return [self.grid("den_i"), self.grid("den_e")]
    @property
    def rho(self): return self.grid("rho")
    @property
    def phi(self): return self.grid("phi")
    @property
    def efx(self): return self.grid("efx")
    @property
    def efy(self): return self.grid("efy")

    def set_species(self, s, x, y, vx, vy):
        x, y, vx, vy = (np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, vx, vy))
        self.L.picsp_ref_species_set(s, _ptr(x), _ptr(y), _ptr(vx), _ptr(vy), len(x))
        self.n[s] = len(x)

    def get_species(self, s):
        n = self.L.picsp_ref_species_count(s)
        out = [np.zeros(n) for _ in range(4)]
        self.L.picsp_ref_species_get(s, *(_ptr(a) for a in out))
        return tuple(out)

    def seed(self, seed): self.L.picsp_ref_seed(seed)
    def rnd(self): return self.L.picsp_ref_rnd()
    def init(self, s, load_type, xdrift=0.0, ydrift=0.0): self.L.picsp_ref_init(s, load_type, xdrift, ydrift)
    def init_both(self, load_type, drift_i=0.0, drift_e=0.0): self.L.picsp_ref_init_both(load_type, drift_i, drift_e)
    def scatterSpecies(self, s): self.L.picsp_ref_scatterSpecies(s)
    def scatterSpeciesVel(self, s): self.L.picsp_ref_scatterSpeciesVel(s)
    def computeRho(self): self.L.picsp_ref_computeRho()
    def solvePotential(self): return bool(self.L.picsp_ref_solvePotential())
    def spectralPotentialSolver(self): return bool(self.L.picsp_ref_spectralPotentialSolver())
    def solve(self): return self.spectralPotentialSolver() if self.solver == 1 else self.solvePotential()
    def computeEF(self): self.L.picsp_ref_computeEF()
    def pushSpecies(self, s): self.L.picsp_ref_pushSpecies(s)
    def rewindSpecies(self, s): self.L.picsp_ref_rewindSpecies(s)
    def computeKE(self, s): return self.L.picsp_ref_computeKE(s)
    def bootstrap(self): self.L.picsp_ref_bootstrap()
    def step(self, nsteps=1, with_dead_vel=False): self.L.picsp_ref_step(nsteps, 1 if with_dead_vel else 0)

    def phase_seconds(self, reset=True):
        out = np.zeros(6)
        self.L.picsp_ref_phase_seconds(_ptr(out), 1 if reset else 0)
        return dict(zip(("deposit", "dead_vel_deposit", "rho", "solve", "ef", "push"), out.tolist()))

    def close(self): self.L.picsp_ref_teardown()

    # -- whole program -------------------------------------------------------
    @classmethod
    def parse_ini(cls, path, lib_path=REF_SO):
        out = np.zeros(20)
        rc = cls.lib(lib_path).picsp_ref_parse_ini(os.fsencode(path), _ptr(out))
        if rc != 0:
            raise RuntimeError("reference parse_ini_file failed")
        keys = ("nTimeSteps timeStep stepSize numxCells numyCells nParticlesI nParticlesE massI massE chargeE "
                "density vthE vthI driftE driftI dumpPeriod solverType loadType ion_spwt electron_spwt").split()
        return dict(zip(keys, out.tolist()))

    @classmethod
    def run_main(cls, ini_path, lib_path=REF_SO):
        """Run the reference's real main() on an ini file; returns {name: ndarray} of everything it wrote."""
        L = cls.lib(lib_path)
        rc = L.picsp_ref_main(os.fsencode(ini_path))
        if rc != 0:
            raise RuntimeError(f"reference main returned {rc}")
        out = {}
        meta = (C.c_longlong * 6)()
        for i in range(L.picsp_ref_h5_count()):
            name = L.picsp_ref_h5_name(i).decode()
            L.picsp_ref_h5_meta(i, meta)
            is_attr, elem, rank, d0, d1, nbytes = list(meta)
            dt = np.float64 if elem == 0 else np.int32
            buf = C.string_at(L.picsp_ref_h5_data(i), nbytes)
            a = np.frombuffer(buf, dtype=dt).copy()
            if rank == 2:
                a = a.reshape(d0, d1)
            out[("@" if is_attr else "") + name] = a
        out["#groups"] = [L.picsp_ref_h5_group_name(i).decode() for i in range(L.picsp_ref_h5_group_count())]
        return out


class WallsChecker:
    """oracle/libpicsp_walls_check.so — the repository's OWN CPU restatement of the PICSP_FLAG_WALLS extension
    (Dirichlet red-black SOR, absorbing walls).  NOT the reference: the reference has no bounded-domain semantics
    (see oracle/walls_check.c).  Same call names as ``Oracle`` where they apply."""

    WALLS_SO = os.path.join(HERE, "libpicsp_walls_check.so")
    TOL, BATCH, MAX_SWEEPS = 1e-12, 16, 100000
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            if not os.path.isfile(cls.WALLS_SO):
                build(ref=False)
            L = C.CDLL(cls.WALLS_SO)
            L.walls_deposit.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, _dp, C.c_long, C.c_double, C.c_int]
            L.walls_rho.argtypes = [C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, C.c_double]
            L.walls_omega.argtypes = [C.c_int, C.c_int]; L.walls_omega.restype = C.c_double
            L.walls_rb_sor.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, C.c_double, C.c_double, C.c_long, C.c_int, C.c_long, _dp]
            L.walls_rb_sor.restype = C.c_long
            L.walls_ef.argtypes = [C.c_int, C.c_int, C.c_double, _dp, _dp, _dp]
            L.walls_push.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, _dp, _dp, _dp, _dp, _dp, _dp, C.c_long,
                                     C.c_double, C.c_double, C.c_int]
            L.walls_push.restype = C.c_long
            cls._lib = L
        return cls._lib

    def __init__(self, numx, numy, dx, dt, mass_i, n_i, n_e, clear=True):
        self.L = self.lib()
        self.numx, self.numy, self.nix, self.niy = int(numx), int(numy), int(numx) + 1, int(numy) + 1
        self.dx, self.dt, self.clear = float(dx), float(dt), bool(clear)
        self.mass, self.charge = [float(mass_i), 1.0], [1.0, -1.0]
        self.spwt = [(1.0 * self.numx * self.numy * self.dx * self.dx) / n for n in (n_i, n_e)]
        nn = self.nix * self.niy
        self.den = [np.zeros(nn), np.zeros(nn)]
        self.rho, self.phi, self.efx, self.efy = (np.zeros(nn) for _ in range(4))
        self.part = [None, None]
        self.omega = self.L.walls_omega(self.numx, self.numy)
        self.last_sweeps, self.last_l2, self.absorbed = 0, 0.0, [0, 0]

    def set_species(self, s, x, y, vx, vy):
        self.part[s] = [np.ascontiguousarray(a, dtype=np.float64).copy() for a in (x, y, vx, vy)]

    def get_species(self, s):
        return tuple(a.copy() for a in self.part[s])

    def grid(self, name):
        return {"den_i": self.den[0], "den_e": self.den[1]}.get(name, getattr(self, name, None))

    def scatterSpecies(self, s):
        x, y = self.part[s][:2]
        self.L.walls_deposit(self.nix, self.niy, self.dx, _ptr(self.den[s]), _ptr(x), _ptr(y), len(x), self.spwt[s], int(self.clear))

    def computeRho(self):
        self.L.walls_rho(self.nix, self.niy, _ptr(self.rho), _ptr(self.den[0]), _ptr(self.den[1]), self.charge[0], self.charge[1])

    def solve(self, fixed_sweeps=0):
        l2 = np.zeros(1)
        self.last_sweeps = self.L.walls_rb_sor(self.nix, self.niy, self.dx, _ptr(self.phi), _ptr(self.rho), self.omega, self.TOL,
                                               self.MAX_SWEEPS, self.BATCH, int(fixed_sweeps), _ptr(l2))
        self.last_l2 = float(l2[0])

    def computeEF(self):
        self.L.walls_ef(self.nix, self.niy, self.dx, _ptr(self.phi), _ptr(self.efx), _ptr(self.efy))

    def _push(self, s, half):
        x, y, vx, vy = self.part[s]
        return self.L.walls_push(self.nix, self.niy, self.dx, self.dt, _ptr(self.efx), _ptr(self.efy), _ptr(x), _ptr(y), _ptr(vx),
                                 _ptr(vy), len(x), self.charge[s], self.mass[s], int(half))

    def pushSpecies(self, s):
        self.absorbed[s] = self._push(s, 0)
        return self.absorbed[s]

    def rewindSpecies(self, s):
        self._push(s, 1)

    def bootstrap(self, sweeps=0):
        self.scatterSpecies(ION); self.scatterSpecies(ELECTRON); self.computeRho(); self.solve(sweeps); self.computeEF()
        self.rewindSpecies(ION); self.rewindSpecies(ELECTRON)

    def step(self, sweeps=0):
        self.scatterSpecies(ION); self.scatterSpecies(ELECTRON); self.computeRho(); self.solve(sweeps); self.computeEF()
        self.pushSpecies(ION); self.pushSpecies(ELECTRON)
