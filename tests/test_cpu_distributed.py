"""World-size-2 gloo test of the multi-rank plumbing (no GPU): the tile sharding map covers every
primary tile exactly once, and the reduce hook sums partial histograms exactly."""
import os

import numpy as np
import torch.multiprocessing as mp

def shard_tile_range(cnt1, cnt2, tstart, ntiles, rank, nranks):
    """Python restatement of k_shard_bounds (corrfunc_b200/csrc/cuda/gridlink.cu): rank r owns the tiles of a contiguous
    range of cells; boundary k is the first cell after the one where the running cost sum(cnt1 * (cnt2 + 1)) reaches
    total * k / nranks."""
    cost = cnt1.astype(np.uint64) * (cnt2.astype(np.uint64) + np.uint64(1))
    after = np.cumsum(cost, dtype=np.uint64)
    before = after - cost
    total = int(after[-1]) if len(after) else 0
    tnext = np.append(tstart[1:], ntiles)
    out = [0, ntiles]
    for side, k in ((0, rank), (1, rank + 1)):
        if k == 0 or k == nranks:
            continue
        t = np.uint64(int(float(total) * (float(k) / float(nranks))))
        hit = np.nonzero((before < t) & (t <= after))[0]
        if len(hit):
            out[side] = int(tnext[hit[0]])
    return out


def test_shard_map_is_a_partition_by_cell():
    """Sharding is by primary cell: every tile of a cell goes to the same rank (the order of the particles inside
    a cell differs between the ranks' replicas, so tiles of one cell must not be split across ranks), every tile is
    owned by exactly one rank (contiguous ranges that tile [0, ntiles)), and the ranks' shares of the cost are balanced."""
    rng = np.random.default_rng(5)
    for ncells in (1, 7, 8, 9, 64, 1000, 17424):
        for clustered in (False, True):
            lam = 114.0 * (np.exp(rng.normal(0.0, 1.0, size=ncells)) if clustered else np.ones(ncells))
            counts = rng.poisson(lam)
            if ncells > 8:
                counts[rng.integers(0, ncells, size=ncells // 8)] = 0  # empty cells hold no tile
            ntile = (counts + 127) // 128
            tstart = np.concatenate(([0], np.cumsum(ntile)[:-1]))
            ntiles = int(ntile.sum())
            tile_cell = np.repeat(np.arange(ncells), ntile)
            for nranks in (1, 2, 3, 4, 8):
                ranges = [shard_tile_range(counts, counts, tstart, ntiles, r, nranks) for r in range(nranks)]
                assert ranges[0][0] == 0 and ranges[-1][1] == ntiles
                for r in range(nranks):
                    lo, hi = ranges[r]
                    assert 0 <= lo <= hi <= ntiles
                    if r + 1 < nranks:
                        assert hi == ranges[r + 1][0]  # the ranges tile [0, ntiles) without gap or overlap
                    # a boundary never falls inside a cell
                    if 0 < lo < ntiles:
                        assert tile_cell[lo] != tile_cell[lo - 1]
                # every boundary sits within one cell of its target: a share misses the mean by at most one cell's cost
                cost = counts.astype(np.float64) * (counts + 1)
                share = np.array([cost[np.unique(tile_cell[lo:hi])].sum() for lo, hi in ranges])
                assert np.abs(share - share.mean()).max() <= cost.max() + 2.0
                if ncells >= 1000 and not clustered:
                    assert share.max() / share.mean() < 1.01


def _worker(rank, world, port, q):
    import torch.distributed as dist

    from corrfunc_b200 import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    npairs = rng.integers(0, 2 ** 40, size=31, dtype=np.uint64)
    ssep = rng.random(31)
    sw = rng.random(31)
    mine = (npairs.copy(), ssep.copy(), sw.copy())
    parallel.make_allreduce(dist)(npairs, ssep, sw)
    q.put((rank, mine, (npairs, ssep, sw)))
    dist.destroy_process_group()


def test_histogram_allreduce_gloo_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tot_n = sum(r[1][0].astype(np.uint64) for r in res)
    tot_s = sum(r[1][1] for r in res)
    tot_w = sum(r[1][2] for r in res)
    for _, _, (n, s, w) in res:
        assert np.array_equal(n, tot_n)  # integer sums exact -> npairs independent of the rank count
        assert np.allclose(s, tot_s, rtol=1e-15) and np.allclose(w, tot_w, rtol=1e-15)


def _worker_replicate(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from corrfunc_b200 import parallel

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a = torch.arange(1001, dtype=torch.float32)  # not a multiple of the world size
    b = torch.arange(7, dtype=torch.float64)
    c = torch.zeros(0, dtype=torch.float32)
    out = parallel.replicate_from_host([a, b, c], torch.device("cpu"), dist)
    q.put((rank, bool(torch.equal(out[0], a) and torch.equal(out[1], b) and out[2].numel() == 0)))
    dist.destroy_process_group()


def test_sharded_upload_allgather_gloo_world2():
    """replicate_from_host: every rank contributes 1/world of the host arrays, all ranks end with the full copy."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_replicate, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(ok for _, ok in res)
