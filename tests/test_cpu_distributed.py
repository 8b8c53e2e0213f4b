"""World-size-2 gloo test of the multi-rank plumbing (no GPU): the tile sharding map covers every
primary tile exactly once, and the reduce hook sums partial histograms exactly."""
import os

import numpy as np
import torch.multiprocessing as mp

SHARD_GROUP = 8  # CFB_SHARD_GROUP in corrfunc_b200/csrc/cuda/cfb_internal.cuh


def owner_of_cell(cell, nranks):
    """Python restatement of cfb_owns_cell (cfb_internal.cuh): rank r owns the cells c with (c / 8) % nranks == r."""
    return (cell // SHARD_GROUP) % nranks if nranks > 1 else 0


def test_shard_map_is_a_partition_by_cell():
    """Sharding is by primary cell: every tile of a cell goes to the same rank (the order of the particles inside
    a cell differs between the ranks' replicas, so tiles of one cell must not be split across ranks), every tile is
    owned by exactly one rank, and the ranks' shares are balanced."""
    rng = np.random.default_rng(5)
    for ncells in (1, 7, 8, 9, 64, 1000, 17424):
        counts = rng.poisson(114, size=ncells)
        ntile = (counts + 127) // 128
        tile_cell = np.repeat(np.arange(ncells), ntile)
        for nranks in (1, 2, 3, 4, 8):
            owner = np.array([owner_of_cell(c, nranks) for c in tile_cell])
            assert owner.min() >= 0 and owner.max() < nranks
            for c in np.unique(tile_cell[:200]):
                assert len(set(owner[tile_cell == c])) == 1
            if ncells >= 1000:
                share = np.bincount(owner, weights=counts[tile_cell] / np.maximum(ntile[tile_cell], 1), minlength=nranks)
                assert share.max() / share.mean() < 1.05


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


def _fast_units(ntiles, tail_first, parts, nrow, rowlen, list_mode):
    """Python restatement of the fast kernel's work units (corrfunc_b200/csrc/cuda/pairs_fast.cu: k_pairs_fast, "Work
    units"): whole tiles up to tail_first, then every tile in `parts` pieces -- one row of neighbour cells each on a box
    lattice (parts = nrow), a share of the neighbour list's rounds of 32 candidates on the RA/DEC lattice."""
    nunits = tail_first + (ntiles - tail_first) * parts
    for gw in range(nunits):
        tile, part, p = gw, 0, 1
        if gw >= tail_first:
            u = gw - tail_first
            tile, part, p = tail_first + u // parts, u % parts, parts
        row_lo, row_hi, base_lo, base_hi = 0, nrow, 0, rowlen
        if p > 1:
            if list_mode:
                nch = (rowlen + 31) >> 5
                base_lo = (part * nch // p) << 5
                base_hi = min(rowlen, ((part + 1) * nch // p) << 5)
            else:
                row_lo, row_hi = part * nrow // p, (part + 1) * nrow // p
        for row in range(row_lo, row_hi):
            for base in range(base_lo, base_hi, 32):
                yield tile, row, base


def test_fast_kernel_work_units_cover_every_round_once():
    """Handing the last tiles out in pieces must not lose or repeat a (tile, row, round of 32 candidates)."""
    for list_mode, nrow, rowlen, parts in ((False, 13, 285, 13), (False, 3, 9, 3), (False, 1, 40, 1), (True, 1, 700, 8),
                                           (True, 1, 33, 8), (True, 1, 0, 8)):
        for ntiles, tail in ((1, 5), (7, 3), (40, 16), (40, 0)):
            tail_first = max(0, ntiles - tail) if parts > 1 else ntiles
            got = sorted(_fast_units(ntiles, tail_first, parts, nrow, rowlen, list_mode))
            want = sorted((t, r, b) for t in range(ntiles) for r in range(nrow) for b in range(0, rowlen, 32))
            assert got == want, (list_mode, nrow, rowlen, parts, ntiles, tail)


def test_two_pass_sort_bucket_layout_and_validity_rule():
    """numpy restatement of the two-pass counting sort's bookkeeping (corrfunc_b200/csrc/cuda/gridlink.cu: k_partition /
    k_place): records of bucket b (2^shift consecutive cells) are written contiguously from start[b << shift]; pass 2
    decides from the position alone whether a slot of the temporary array holds a record -- the last bucket whose base is
    <= p, and p - base < count -- also when buckets are empty or share a base.  Every particle must land in its cell's
    run of the final array exactly once."""
    rng = np.random.default_rng(11)
    PAD, MAXB = 4, 512
    for ncells, n in ((1, 10), (5, 0), (700, 3000), (5000, 20000), (5000, 300)):
        cidx = rng.integers(0, ncells, size=n)
        if ncells >= 700:
            cidx[cidx % 7 == 3] = 0  # runs of empty cells and empty buckets
        count = np.bincount(cidx, minlength=ncells)
        padded = (count + PAD - 1) // PAD * PAD
        start = np.concatenate(([0], np.cumsum(padded)[:-1]))
        npad = int(padded.sum())
        shift = 0
        while ((ncells + (1 << shift) - 1) >> shift) > MAXB:
            shift += 1
        nb = (ncells + (1 << shift) - 1) >> shift
        base = start[(np.arange(nb) << shift)]
        # pass 1: arrival order within the bucket
        rec = np.full(npad, -1)
        bcur = np.zeros(nb, dtype=np.int64)
        for i in rng.permutation(n):
            b = cidx[i] >> shift
            rec[base[b] + bcur[b]] = i
            bcur[b] += 1
        # pass 2: validity from the position
        cur = start.copy()
        final = np.full(max(npad, 1), -1)
        for p in range(npad):
            lo, hi = 0, nb
            while hi - lo > 1:
                mid = (lo + hi) >> 1
                if base[mid] <= p:
                    lo = mid
                else:
                    hi = mid
            valid = p - base[lo] < bcur[lo]
            assert valid == (rec[p] >= 0), (ncells, n, p)
            if valid:
                c = cidx[rec[p]]
                final[cur[c]] = rec[p]
                cur[c] += 1
        for c in range(ncells):
            run = final[start[c]:start[c] + count[c]]
            assert np.all(run >= 0) and np.all(cidx[run] == c)
        assert np.count_nonzero(final >= 0) == n
