"""GPU, 2 ranks over NCCL: the sharded path equals the single-GPU path (skipped with < 2 GPUs)."""
import os
import socket

import numpy as np
import pytest

from oracle.oracle import ELECTRON, ION, Oracle, normalise
from tests.helpers import GRIDS, assert_grid_close, relerr

pytestmark = pytest.mark.gpu


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, numx, n, solver, out_dir, flags=0):
    import torch
    import torch.distributed as dist
    from picsp_b200 import Params, Simulation
    from picsp_b200.sim import shard_range
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    nm = normalise()
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    o.seed(8); o.init(ION, 1); o.init(ELECTRON, 1)
    lo, hi = shard_range(n, rank, world)
    sim = Simulation(Params(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, solverType=solver, device=rank,
                            capacity=(hi - lo, hi - lo), flags=flags))
    for s in (ION, ELECTRON):
        sim.set_species(s, *(a[lo:hi] for a in o.get_species(s)))
    uid = [Simulation.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim.comm_attach(uid[0], rank, world)
    sim.bootstrap(); sim.step(3)
    out = {g: sim.grid(g) for g in ("rho", "phi", "efx", "efy", "den_i", "den_e")}
    out["peer"] = np.array([int(sim.comm_peer_reduction())])   # den_*: collective, global sum on every rank
    d = sim.dump(root=(rank == 0))           # collective: den.i / den.e reduced to rank 0, phi on rank 0, KE global
    for k, v in d.items():
        out["dump_" + k] = v
    for s, nm_ in ((ION, "i"), (ELECTRON, "e")):
        out["part_" + nm_] = np.stack(sim.get_species(s))
        out["ke_" + nm_] = np.array([sim.computeKE(s)])
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), **out)
    sim.close()
    dist.barrier(); dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("solver,flags", [(1, 0), (2, 0), (1, 128)], ids=["spectral-peer", "sor-peer", "spectral-nccl-only"])
def test_two_ranks_equal_oracle(tmp_path, solver, flags):
    """flags 0: the partial rho is summed by the library's own peer-memory kernels (when the two GPUs can map each other);
    128 = PICSP_FLAG_NCCL_ONLY: by ncclAllReduce.  Both against the single-rank oracle."""
    import torch.multiprocessing as mp
    from picsp_b200.sim import shard_range
    numx, n, world = 64, 40_000, 2
    mp.spawn(_worker, args=(world, _free_port(), numx, n, solver, str(tmp_path), flags), nprocs=world, join=True)
    nm = normalise()
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    o.seed(8); o.init(ION, 1); o.init(ELECTRON, 1)
    o.bootstrap(); o.step(3)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
    assert res[0]["peer"][0] == res[1]["peer"][0] and (flags == 0 or res[0]["peer"][0] == 0)
    print("rho reduction:", "own peer-memory kernels" if res[0]["peer"][0] else "ncclAllReduce")
    for g in ("rho", "phi", "efx", "efy"):
        assert np.array_equal(res[0][g], res[1][g]), f"{g} differs between ranks (redundant solve must be bit-identical)"
        assert relerr(res[0][g], o.grid(g)) < 1e-11, g
    # what a sharded run dumps (writeSpecies, main.cpp:1166-1172) is the density of ALL particles: the download is a
    # collective that sums the per-rank partial densities, the asynchronous dump reduces them to rank 0
    nix = numx + 1
    for g in ("den_i", "den_e"):
        assert np.array_equal(res[0][g], res[1][g]), f"{g}: the two ranks disagree on the global density"
        assert_grid_close(res[0][g], o.grid(g), nix, nix, 1e-12, g)
        assert np.array_equal(res[0]["dump_" + g], res[0][g]), f"dump {g} on rank 0 is not the global density"
        assert "dump_" + g not in res[1].files
    assert np.array_equal(res[0]["dump_phi"], res[0]["phi"])
    for r in range(world):
        assert np.array_equal(res[r]["dump_rows_e"].T, res[r]["part_e"]) and np.array_equal(res[r]["dump_rows_i"].T, res[r]["part_i"])
        assert res[r]["dump_ke"][1] == res[r]["ke_e"][0] or abs(res[r]["dump_ke"][1] - res[r]["ke_e"][0]) <= 1e-13 * abs(res[r]["ke_e"][0])
    for s, nm_ in ((ION, "i"), (ELECTRON, "e")):
        want = np.stack(o.get_species(s))
        got = np.concatenate([res[r]["part_" + nm_] for r in range(world)], axis=1)
        for k in range(4):
            assert relerr(got[k], want[k]) < 1e-11
        assert abs(res[0]["ke_" + nm_][0] - o.computeKE(s)) <= 1e-11 * abs(o.computeKE(s))
        assert res[0]["ke_" + nm_][0] == res[1]["ke_" + nm_][0]


@pytest.mark.skipif(_ngpu() < 2, reason="needs 2 GPUs")
def test_host_program_sharded_over_two_gpus_writes_the_single_rank_file(tmp_path):
    """`picsp_b200_run input.ini` launched once per GPU (RANK / WORLD_SIZE / LOCAL_RANK) on BASELINE config 1: ONE HDF5
    file, rows of both ranks in list order, den.i / den.e reduced over the ranks.  Equal to the single-GPU
    file (den to 1e-13) and to the reference's own main() (golden)."""
    import subprocess
    from picsp_b200 import host
    from picsp_b200.lib import PKG
    from tests import h5mini
    from tests.helpers import GOLDEN, load_golden
    ini = os.path.join(GOLDEN, "input_ini_shipped.ini")
    exe = os.path.join(PKG, "picsp_b200_run")
    out2, out1 = str(tmp_path / "two.h5"), str(tmp_path / "one.h5")
    procs = [subprocess.Popen([exe, ini, "--out", out2, "--steps", "100"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                              env=dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r))) for r in range(2)]
    try:
        outs = [p.communicate(timeout=240) for p in procs]
    finally:
        for p in procs:                 # a rank that died must not leave the other one waiting in a collective
            if p.poll() is None:
                p.kill()
    assert all(p.returncode == 0 for p in procs), [o[1][-1500:] for o in outs]
    assert "TS: 100 \t delta_phi:" in outs[0][0] and outs[1][0] == "", "rank 0 alone prints the reference's lines"
    assert not os.path.exists(out2 + ".ncclid")
    host.run(ini, out1, max_steps=100, quiet=True)
    a, b, g = h5mini.File(out2), h5mini.File(out1), load_golden("whole_run_input_ini")
    assert a.attrs().keys() == b.attrs().keys() and sorted(a.groups()) == sorted(b.groups())
    worst = 0.0
    for grp in ("particle.i", "particle.e", "den.i", "den.e", "phi"):
        assert a.datasets(grp) == b.datasets(grp) == sorted(["0", "50", "100"])
        for ts in ("0", "50", "100"):
            x, y = a.read(f"/{grp}/{ts}"), b.read(f"/{grp}/{ts}")
            assert x.shape == y.shape
            if grp.startswith("den"):
                x, y = x[1:-1, 1:-1], y[1:-1, 1:-1]
            e = relerr(x, y)
            tol = 1e-13 if grp.startswith("den") else (1e-9 if (grp, ts) == ("phi", "0") else 1e-11)
            assert e <= tol, (grp, ts, e)
            worst = max(worst, e)
    assert np.allclose(a.read("/timedata/energy"), b.read("/timedata/energy"), rtol=1e-12, atol=0)
    assert relerr(a.read("/particle.e/0"), g["particle_e_0"]) <= 1e-12 and relerr(a.read("/den.i/50")[1:-1, 1:-1], g["den_i_50"][1:-1, 1:-1]) <= 1e-12
    print(f"2-rank host run vs 1-rank: worst rel err {worst:.2e}")
