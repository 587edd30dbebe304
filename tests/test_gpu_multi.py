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


def _worker(rank, world, port, numx, n, solver, out_dir):
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
                            capacity=(hi - lo, hi - lo)))
    for s in (ION, ELECTRON):
        sim.set_species(s, *(a[lo:hi] for a in o.get_species(s)))
    uid = [Simulation.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    sim.comm_attach(uid[0], rank, world)
    sim.bootstrap(); sim.step(3)
    out = {g: sim.grid(g) for g in ("rho", "phi", "efx", "efy", "den_i", "den_e")}   # den_*: collective, global sum on every rank
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
@pytest.mark.parametrize("solver", [1, 2])
def test_two_ranks_equal_oracle(tmp_path, solver):
    import torch.multiprocessing as mp
    from picsp_b200.sim import shard_range
    numx, n, world = 64, 40_000, 2
    mp.spawn(_worker, args=(world, _free_port(), numx, n, solver, str(tmp_path)), nprocs=world, join=True)
    nm = normalise()
    o = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=solver)
    o.seed(8); o.init(ION, 1); o.init(ELECTRON, 1)
    o.bootstrap(); o.step(3)
    res = [np.load(tmp_path / f"r{r}.npz") for r in range(world)]
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
