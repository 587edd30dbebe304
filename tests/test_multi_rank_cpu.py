"""CPU, world_size 2 over gloo: the host-side logic of the multi-GPU path.

The CUDA path shards particles by index range, lets every rank deposit a full partial grid and
sums the partial charge densities with one all-reduce.  Here the same decomposition
is exercised on CPU with the oracle as the per-rank worker and gloo as the collective: the
all-reduced partial rho must equal the single-rank rho, and the sharding must be a partition."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle.oracle import ELECTRON, ION, Oracle, normalise
from picsp_b200.sim import shard_range


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def test_shard_range_is_a_partition():
    for n in (0, 1, 7, 10_000, 500_000_000, 2_000_000_001):
        for world in (1, 2, 3, 4, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            for a, b in zip(cuts, cuts[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, numx, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nm = normalise()
    full = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=2)
    full.seed(4); full.init(ION, 1); full.init(ELECTRON, 1)
    # this rank's shard, with the GLOBAL specific weight (spwt is defined on the global count)
    part = Oracle(numx, numx, nm["dx"], nm["dt"], nm["mass_i"], n, n, vth_i=nm["vth_i"], solver=2)
    for s in (ION, ELECTRON):
        lo, hi = shard_range(n, rank, world)
        part.set_species(s, *(a[lo:hi] for a in full.get_species(s)))
        part.spwt[s] = full.spwt[s]
    for _ in range(2):   # two accumulating deposits: per-rank accumulation and fold commute with the sum
        part.scatterSpecies(ION); part.scatterSpecies(ELECTRON)
    part.computeRho()
    rho = torch.from_numpy(part.rho.copy())
    dist.all_reduce(rho)                      # the one exchange step of the path
    for _ in range(2):
        full.scatterSpecies(ION); full.scatterSpecies(ELECTRON)
    full.computeRho()
    err = np.abs(rho.numpy() - full.rho).max() / np.abs(full.rho).max()
    # dump semantics (picsp_dump_begin): the per-rank partial densities reduced to rank 0 are the density
    # the single-rank run dumps, edge nodes (folded twice) included
    den_err = 0.0
    for s in (ION, ELECTRON):
        d = torch.from_numpy(part.den[s].copy())
        dist.reduce(d, dst=0)
        if rank == 0:
            den_err = max(den_err, float(np.abs(d.numpy() - full.den[s]).max() / np.abs(full.den[s]).max()))
    # every rank then solves redundantly on identical input
    part.rho[...] = rho.numpy(); part.solve(); part.computeEF()
    phi = torch.from_numpy(part.phi.copy())
    ref = phi.clone(); dist.broadcast(ref, src=0)
    same_phi = bool(torch.equal(phi, ref))
    np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([err, float(same_phi), den_err]))
    dist.destroy_process_group()


def test_partial_rho_allreduce_equals_single_rank(tmp_path):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, 24, 4001, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        err, same, den_err = np.load(tmp_path / f"r{r}.npy")
        assert err < 1e-13, err
        assert same == 1.0
        assert den_err < 1e-13, den_err
