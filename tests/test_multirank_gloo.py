"""World-size-2 checks (gloo, CPU) of the host-side logic of the sharded SMC: block partition, the collective
protocol (all-gather of per-rank canonical block partials, replicated weight algebra) and index-keyed random numbers
give results that are bit-identical to a single rank.  The device implementation follows the same protocol with NCCL
(scripts/smc_multigpu.py --check verifies it on GPUs)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import rng as orng
from oracle import smc as osmc


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, n, npar, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynare_jl_b200.samplers import shard_range
    first, nl = shard_range(n, rank, world)
    g = np.random.default_rng(7)                         # every rank draws the same global system, keeps its slice
    w_prev = g.random(n); w_prev /= w_prev.sum()
    loglh = g.normal(-700, 5, n); para = g.normal(size=(n, npar))
    sl = slice(first, first + nl)
    # correction: local incremental weights, all-gather, then the replicated canonical algebra
    wt_l = w_prev[sl] * np.exp(0.01 * loglh[sl])
    box = [None] * world
    dist.all_gather_object(box, wt_l)
    wt = np.concatenate(box)
    zhat = osmc.canonical_sum(wt); w = wt / zhat; ess = 1.0 / osmc.canonical_sum(w * w)
    # selection: own output slots from the replicated cumulative weights and index-keyed uniforms
    u = orng.uniform(5, np.arange(first, first + nl), 3, orng.STREAM_RESAMPLE)
    idx_l, _ = osmc.resample_indices(w, u)
    # moments: per-rank block partials (blocks never straddle ranks), all-gather, sequential sum over blocks
    part_l = np.array([[osmc.canonical_sum((w[sl] * para[sl, p])[b * 1024:(b + 1) * 1024]) for p in range(npar)]
                       for b in range(nl // 1024)])
    dist.all_gather_object(box, part_l)
    part = np.concatenate(box)
    mu = np.zeros(npar)
    for b in range(part.shape[0]):
        mu = mu + part[b]
    box2 = [None] * world
    dist.all_gather_object(box2, idx_l)
    if rank == 0:
        q.put(dict(zhat=zhat, ess=ess, idx=np.concatenate(box2), mu=mu, w=w))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sharded_smc_protocol_matches_single_rank():
    n, npar, world = 4096, 3, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, npar, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    g = np.random.default_rng(7)
    w_prev = g.random(n); w_prev /= w_prev.sum()
    loglh = g.normal(-700, 5, n); para = g.normal(size=(n, npar))
    w, zhat, ess = osmc.reweight(w_prev, loglh, 0.01)
    idx, _ = osmc.resample_indices(w, orng.uniform(5, np.arange(n), 3, orng.STREAM_RESAMPLE))
    mu, _ = osmc.moments(para, w)
    assert got["zhat"] == zhat and got["ess"] == ess and np.array_equal(got["w"], w)     # bit-identical
    assert np.array_equal(got["idx"], idx)
    assert np.array_equal(got["mu"], mu)


def test_shard_range_partition():
    from dynare_jl_b200.samplers import shard_range
    assert [shard_range(16384, r, 8) for r in range(8)] == [(2048 * r, 2048) for r in range(8)]
    with pytest.raises(ValueError):
        shard_range(1000, 0, 3)
