#!/usr/bin/env python
"""bench.py -- pair-counting throughput on B200 (one JSON line on stdout, rank 0).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c5|c1|c2|c3|c4] [--impl ours|reference]

A "step" is one complete pair count of the workload (gridlink + cell-pair enumeration + pair kernel +
histogram read-back).  Workloads are the BASELINE.json configs with the synthetic generators of
SURVEY.md section 8(d); the default (headline) is c5: xi on 100M uniform points in a 2 Gpc/h periodic
box, 30 log bins 0.1-150 Mpc/h, float32.

  value  = reference-equivalent candidate pair evaluations per second with inputs resident in HBM:
           N_cand = sum over the REFERENCE's cell pairs of N1*N2 (same cell: N(N-1)/2) -- a property
           of the workload, identical for both arms, so the ratio of the arms is a wall-time ratio.
  e2e    = the same with HOST (pinned) buffers through the C ABI: H2D + gridlink + pairs + D2H timed.
  roofline = the pair kernel alone: separations actually computed by the kernel (n_eval, counted on
           the device) x 8 FLOP / kernel time, against the FP32 (or FP64) ALU peak.

--impl reference times the reference's own CPU implementation (oracle/_ref, OpenMP, all host cores)
on a bounded random subsample of the same workload (same box, same bins).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (stat, N, L, bins, dtype, extra)
    "c1": dict(stat="DD", N=1_200_000, L=420.0, bins=("log", 0.1, 25.0, 15), dtype="f64", seed=1001),
    "c2": dict(stat="wp", N=1_200_000, L=420.0, bins=("log", 0.1, 25.0, 15), dtype="f64", seed=1001, pimax=40.0),
    "c2rppi": dict(stat="DDrppi", N=1_200_000, L=420.0, bins=("log", 0.1, 25.0, 15), dtype="f64", seed=1001, pimax=40.0),
    "c2rppi32": dict(stat="DDrppi", N=1_200_000, L=420.0, bins=("log", 0.1, 25.0, 15), dtype="f32", seed=1001, pimax=40.0),
    "c2wp32": dict(stat="wp", N=1_200_000, L=420.0, bins=("log", 0.1, 25.0, 15), dtype="f32", seed=1001, pimax=40.0),
    "c3": dict(stat="DDsmu", N=10_000_000, L=1000.0, bins=("log", 0.1, 50.0, 21), dtype="f64", seed=1003,
               mu_max=1.0, nmu=20, weights=True, avg=True),
    "c4": dict(stat="DDtheta", N=2_000_000, L=0.0, bins=("log", 0.01, 10.0, 21), dtype="f64", seed=1004),
    # float32 twins of configs 1, 3 and 4: parity cases only (full-size goldens), not bench lines
    "c1f32": dict(stat="DD", N=1_200_000, L=420.0, bins=("log", 0.1, 25.0, 15), dtype="f32", seed=1001),
    "c3f32": dict(stat="DDsmu", N=10_000_000, L=1000.0, bins=("log", 0.1, 50.0, 21), dtype="f32", seed=1003,
                  mu_max=1.0, nmu=20, weights=True, avg=True),
    "c4f32": dict(stat="DDtheta", N=2_000_000, L=0.0, bins=("log", 0.01, 10.0, 21), dtype="f32", seed=1004),
    # survey geometry (full-sky shell 500 < D < 1500, uniform in volume), comoving distances
    "m1": dict(stat="DDrppi_mocks", N=4_000_000, L=0.0, bins=("log", 0.1, 25.0, 15), dtype="f64", seed=1011, pimax=40.0),
    "m2": dict(stat="DDsmu_mocks", N=4_000_000, L=0.0, bins=("log", 0.1, 50.0, 21), dtype="f64", seed=1012,
               mu_max=1.0, nmu=20, weights=True, avg=True),
    "c5d": dict(stat="xi", N=100_000_000, L=2000.0, bins=("log", 0.1, 150.0, 31), dtype="f64", seed=1006),
    "c5": dict(stat="xi", N=100_000_000, L=2000.0, bins=("log", 0.1, 150.0, 31), dtype="f32", seed=1006),
    # config 5 is "xi AND DD(autocorr=1, periodic)": the same points through countpairs(), whose lattice
    # spans the data extent instead of [0, L]
    "c5DD": dict(stat="DD", N=100_000_000, L=2000.0, bins=("log", 0.1, 150.0, 31), dtype="f32", seed=1006),
}
# mocks: up to the first range test of a pair -- 3 sub + 3 add (perp, par), 2 mul + add + fma (s.l), its square,
# mul + 2 fma (s^2): 14 lane-instructions, 17 FLOP
FLOP_PER_EVAL = {"DD": 8, "xi": 8, "wp": 6, "DDrppi": 7, "DDsmu": 9, "DDtheta": 10, "DDrppi_mocks": 17, "DDsmu_mocks": 17}
INSTR_PER_EVAL = {"DD": 6, "xi": 6, "wp": 5, "DDrppi": 6, "DDsmu": 7, "DDtheta": 8, "DDrppi_mocks": 14, "DDsmu_mocks": 14}
HOST_ONLY = ("DDtheta", "DDrppi_mocks", "DDsmu_mocks")  # the host layer converts the angles itself: host buffers only


def config_by_name(name):
    """CONFIGS[name], or '<cfg>sd<N>M': that config at the same number density with N million points
    (what --npart N --same-density runs), e.g. c5sd10M."""
    if name in CONFIGS:
        return dict(CONFIGS[name])
    base, _, n = name.partition("sd")
    cfg = dict(CONFIGS[base])
    npart = int(float(n.rstrip("M")) * 1e6)
    cfg["L"] = float(cfg["L"] * (npart / cfg["N"]) ** (1.0 / 3.0))
    cfg["N"] = npart
    return cfg


def make_bins(spec):
    kind, lo, hi, n = spec
    return np.logspace(np.log10(lo), np.log10(hi), n)


def gen_points(cfg, n, dtype):
    rng = np.random.default_rng(cfg["seed"])
    if cfg["stat"] == "DDtheta":
        ra = (360.0 * rng.random(n)).astype(dtype)
        dec = np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0)).astype(dtype)
        rng2 = np.random.default_rng(cfg["seed"] + 1)
        ra2 = (360.0 * rng2.random(n)).astype(dtype)
        dec2 = np.degrees(np.arcsin(2.0 * rng2.random(n) - 1.0)).astype(dtype)
        return dict(ra=ra, dec=dec, ra2=ra2, dec2=dec2)
    if cfg["stat"].endswith("_mocks"):
        out = dict(ra=(360.0 * rng.random(n)).astype(dtype),
                   dec=np.degrees(np.arcsin(2.0 * rng.random(n) - 1.0)).astype(dtype),
                   d=np.cbrt(500.0 ** 3 + (1500.0 ** 3 - 500.0 ** 3) * rng.random(n)).astype(dtype))
        if cfg.get("weights"):
            out["w"] = (1.0 - rng.random(n)).astype(dtype)
        return out
    out = {}
    for k in "xyz":  # one axis at a time keeps the peak host memory at ~2 arrays
        out[k] = (rng.random(n) * cfg["L"]).astype(dtype)
    if cfg.get("weights"):
        out["w"] = (1.0 - rng.random(n)).astype(dtype)
    return out


def n_cand_box(counts, refine, periodic=True):
    """Candidate pairs of the reference's cell-pair set (generate_cell_pairs, no min-sep pruning):
    unordered pairs of particles whose reference cells are within +-refine of each other."""
    c = counts.astype(np.float64)
    s = np.zeros_like(c)
    rx, ry, rz = refine
    seen = set()
    for dx in range(-rx, rx + 1):
        for dy in range(-ry, ry + 1):
            for dz in range(-rz, rz + 1):
                key = (dx % c.shape[0], dy % c.shape[1], dz % c.shape[2])
                if key in seen:  # tiny lattices: the same neighbour reached twice
                    continue
                seen.add(key)
                s += np.roll(c, shift=(dx, dy, dz), axis=(0, 1, 2))
    return float((c * s).sum() - c.sum()) / 2.0


def n_cand_open(counts, refine):
    """n_cand_box for a lattice without periodic wrap (zero padding instead of roll)."""
    c = counts.astype(np.float64)
    rx, ry, rz = refine
    pad = np.pad(c, ((rx, rx), (ry, ry), (rz, rz)))
    s = np.zeros_like(c)
    for dx in range(2 * rx + 1):
        for dy in range(2 * ry + 1):
            for dz in range(2 * rz + 1):
                s += pad[dx:dx + c.shape[0], dy:dy + c.shape[1], dz:dz + c.shape[2]]
    return float((c * s).sum() - c.sum()) / 2.0


def n_cand_data_extent(cfg, pts, n, bins):
    """Candidate pairs of the reference's lattice over the DATA extent, for the statistics whose lattice is not the
    [0,L]^3 one: DDsmu (periodic; countpairs_s_mu_impl.c.src:190-234, no 0.05 rule, boost is a no-op) and the mocks
    statistics (open; countpairs_rp_pi_mocks_impl.c.src:404-470).  Computed in double: a throughput denominator."""
    stat = cfg["stat"]
    rmax = float(bins[-1])
    if stat.endswith("_mocks"):
        ra, dec, d = (np.asarray(pts[k][:n], dtype=np.float64) for k in ("ra", "dec", "d"))
        xyz = [d * np.cos(np.radians(dec)) * np.cos(np.radians(ra)), d * np.cos(np.radians(dec)) * np.sin(np.radians(ra)),
               d * np.sin(np.radians(dec))]
        sep = np.sqrt(rmax ** 2 + cfg["pimax"] ** 2) if stat == "DDrppi_mocks" else rmax
        maxsize = [sep, sep, sep]
    else:
        xyz = [np.asarray(pts[k][:n], dtype=np.float64) for k in "xyz"]
        maxsize = [rmax, rmax, rmax * cfg["mu_max"]]
    lo = [float(a.min()) for a in xyz]
    ext = [float(a.max()) - l for a, l in zip(xyz, lo)]
    rf = [2, 2, 1]
    if stat.endswith("_mocks"):
        rf = [1 if maxsize[0] < 0.05 * e else r for r, e in zip(rf, ext)]

    def mesh(rf):
        return [max(2, min(100, int(r * e / m))) for r, e, m in zip(rf, ext, maxsize)]

    nm = mesh(rf)
    if stat.endswith("_mocks") and (max(nm) <= 10 or n / (nm[0] * nm[1] * nm[2]) >= 250) and max(nm) < 100:
        rf = [rf[0] + 1, rf[1] + 1, rf[2]]
        nm = mesh(rf)
    idx = [np.minimum((((a - l) * (m / e)).astype(np.int64)), m - 1) for a, l, m, e in zip(xyz, lo, nm, ext)]
    lin = (idx[0] * nm[1] + idx[1]) * nm[2] + idx[2]
    counts = np.bincount(lin, minlength=nm[0] * nm[1] * nm[2]).reshape(nm)
    return n_cand_open(counts, rf) if stat.endswith("_mocks") else n_cand_box(counts, rf)


def ref_cell_counts(pts, L, nmesh, dtype):
    """Reference cell occupancies for [0,L]^3 lattices (xi/wp): ix=(int)(x*xinv), clamp."""
    idx = []
    for k, n in zip("xyz", nmesh):
        inv = dtype(1.0 / (dtype(L) / dtype(n)))
        i = (pts[k] * inv).astype(np.int64)
        np.minimum(i, n - 1, out=i)
        idx.append(i)
    lin = (idx[0] * nmesh[1] + idx[1]) * nmesh[2] + idx[2]
    return np.bincount(lin, minlength=nmesh[0] * nmesh[1] * nmesh[2]).reshape(nmesh)


def workload_config(args, cfg, N, input_bytes):
    """The workload both arms are measured on, as a dict that depends on nothing but the workload itself (the driver
    compares the two arms' `config` for equality)."""
    return {"workload": "%s %s: N=%d L=%g bins=%s%s" % (args.config, cfg["stat"], N, cfg["L"], cfg["bins"],
                                                        "" if not args.npart else " (REDUCED N, not the headline size)"),
            "dtype": cfg["dtype"],
            "l2": "inputs larger than L2" if input_bytes >= 256 * 1024 * 1024 else "256 MiB L2 flush between iterations"}


def input_bytes_of(cfg, N):
    per = {"DDtheta": 4, "DDrppi_mocks": 3, "DDsmu_mocks": 3}.get(cfg["stat"], 3) + (1 if cfg.get("weights") else 0)
    return per * N * (4 if cfg["dtype"] == "f32" else 8)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([t.strip() for t in line.split(",")])

    def stop(self):
        if not self.proc:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def run_ours(args, cfg):
    import torch
    import torch.distributed as dist

    from corrfunc_b200 import _capi, _lib, parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    os.environ["CORRFUNC_B200_DEVICE"] = str(local)
    if args.inlib and world == 1:
        # ONE process, ONE C call, args.gpus devices: the sharding happens inside the library (CORRFUNC_B200_NGPUS)
        os.environ["CORRFUNC_B200_NGPUS"] = str(args.gpus)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    if world > 1:
        parallel.enable_distributed(dist, dev)
    if args.occ:
        lib.cfb_set_target_occupancy(args.occ)

    stat = cfg["stat"]
    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    tdtype = torch.float32 if dtype == np.float32 else torch.float64
    N = args.npart or cfg["N"]
    bins = make_bins(cfg["bins"])
    pts = gen_points(cfg, N, dtype)
    keys = list(pts.keys())
    # host copies in pinned memory (e2e) and device-resident copies (value)
    pinned = {k: torch.from_numpy(pts[k]).pin_memory() for k in keys}
    resident = {k: pinned[k].to(dev) for k in keys}
    opt_kw = dict(periodic=True, need_avg_sep=bool(cfg.get("avg")), boxsize=cfg["L"] if cfg["L"] > 0 else None,
                  is_comoving_dist=stat.endswith("_mocks"))
    wtype = "pair_product" if cfg.get("weights") else None
    _capi._declare(lib)

    def one_call(bufs, bf):
        """One pass through the C ABI; bufs: dict of torch tensors (host pinned or device)."""
        o = _capi.default_options(dtype, **opt_kw)
        P = {k: C.c_void_p(v.data_ptr()) for k, v in bufs.items()}
        if wtype:
            # the epilogue needs the weights on the host (self-pair term) -> always the pinned copy
            e, keep = _capi.make_extra(pinned["w"].numpy(), pinned["w"].numpy(), wtype, dtype)
        else:
            e, keep = _capi.make_extra(None, None, None, dtype)
        if stat == "xi":
            r = _capi.ResultsXi()
            st = lib.countpairs_xi(N, P["x"], P["y"], P["z"], cfg["L"], 1, bf, C.byref(r), C.byref(o), C.byref(e))
            free = lib.free_results_xi
        elif stat == "DD":
            r = _capi.ResultsDD()
            st = lib.countpairs(N, P["x"], P["y"], P["z"], N, P["x"], P["y"], P["z"], 1, 1, bf, C.byref(r), C.byref(o), C.byref(e))
            free = lib.free_results
        elif stat == "wp":
            r = _capi.ResultsWp()
            st = lib.countpairs_wp(N, P["x"], P["y"], P["z"], cfg["L"], 1, bf, cfg["pimax"], C.byref(r), C.byref(o), C.byref(e))
            free = lib.free_results_wp
        elif stat == "DDrppi":
            r = _capi.ResultsRpPi()
            st = lib.countpairs_rp_pi(N, P["x"], P["y"], P["z"], N, P["x"], P["y"], P["z"], 1, 1, bf, cfg["pimax"],
                                      C.byref(r), C.byref(o), C.byref(e))
            free = lib.free_results_rp_pi
        elif stat == "DDsmu":
            r = _capi.ResultsSMu()
            st = lib.countpairs_s_mu(N, P["x"], P["y"], P["z"], N, P["x"], P["y"], P["z"], 1, 1, bf, cfg["mu_max"],
                                     cfg["nmu"], C.byref(r), C.byref(o), C.byref(e))
            free = lib.free_results_s_mu
        elif stat == "DDtheta":
            # RA/DEC -> unit vectors happens on the host (glibc trig, for bit parity): host buffers only
            r = _capi.ResultsTheta()
            st = lib.countpairs_theta_mocks(N, C.c_void_p(pinned["ra"].data_ptr()), C.c_void_p(pinned["dec"].data_ptr()),
                                            N, C.c_void_p(pinned["ra2"].data_ptr()), C.c_void_p(pinned["dec2"].data_ptr()),
                                            1, 0, bf, C.byref(r), C.byref(o), C.byref(e))
            free = lib.free_results_countpairs_theta
        elif stat in ("DDrppi_mocks", "DDsmu_mocks"):
            # RA/DEC/distance -> Cartesian happens on the host (glibc trig, for bit parity): host buffers only
            H3 = [C.c_void_p(pinned[k].data_ptr()) for k in ("ra", "dec", "d")]
            if stat == "DDrppi_mocks":
                r = _capi.ResultsMocksRpPi()
                st = lib.countpairs_mocks(N, H3[0], H3[1], H3[2], N, H3[0], H3[1], H3[2], 1, 1, bf, cfg["pimax"], 1,
                                          C.byref(r), C.byref(o), C.byref(e))
                free = lib.free_results_mocks
            else:
                r = _capi.ResultsMocksSMu()
                st = lib.countpairs_mocks_s_mu(N, H3[0], H3[1], H3[2], N, H3[0], H3[1], H3[2], 1, 1, bf, cfg["mu_max"],
                                               cfg["nmu"], 1, C.byref(r), C.byref(o), C.byref(e))
                free = lib.free_results_mocks_s_mu
        else:
            raise ValueError(stat)
        if st != 0:
            raise RuntimeError("C call failed: %s" % lib.cfb_last_error())
        nb = r.nbin if hasattr(r, "nbin") else r.nsbin
        tot = int(np.ctypeslib.as_array(r.npairs, shape=(nb,)).astype(np.uint64)[1:].sum()) if stat not in ("DDsmu", "DDrppi", "DDrppi_mocks", "DDsmu_mocks") else -1
        free(C.byref(r))
        return tot

    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    input_bytes = sum(v.numel() * v.element_size() for v in resident.values())
    need_flush = input_bytes < 256 * 1024 * 1024

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    with _capi.binfile_for(bins) as bf:
        # ---- warm-up ----
        for _ in range(max(args.warmup, 3)):
            if need_flush:
                flush.zero_()
            tot = one_call(resident, bf)
        st0 = _lib.last_stats()
        # ---- timed: inputs resident in HBM ----
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        kern_ms, grid_ms, dev_ms, n_eval, launches = [], [], [], 0, 0
        barrier()
        t_acc = 0.0
        for _ in range(args.steps):
            if need_flush:
                flush.zero_()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            one_call(resident, bf)  # synchronous: returns after the histogram has been read back
            t_acc += time.perf_counter() - t0
            s = _lib.last_stats()
            kern_ms.append(s["ms_pairs"])
            grid_ms.append(s["ms_gridlink"])
            dev_ms.append(s["ms_total_device"])
            n_eval = s["n_eval"]
            launches += s["kernel_launches"]
        barrier()
        t_res = t_acc / args.steps
        # ---- timed: end to end from pinned host buffers ----
        if cfg["N"] <= 20_000_000:
            one_call(pinned, bf) if not (world > 1 and stat not in HOST_ONLY) else None  # untimed: first touch of the host path
        barrier()
        t_acc = 0.0
        for _ in range(args.steps):
            if need_flush:
                flush.zero_()
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            if world > 1 and stat not in HOST_ONLY:
                # each rank uploads 1/world of the host arrays, one NVLink all-gather completes the replicas
                dev_t = parallel.replicate_from_host([pinned[k] for k in keys], dev, dist)
                torch.cuda.current_stream().synchronize()
                one_call(dict(zip(keys, dev_t)), bf)
            else:
                one_call(pinned, bf)
            t_acc += time.perf_counter() - t0
        barrier()
        t_e2e = t_acc / args.steps
        st1 = _lib.last_stats()
        # ---- end to end from PAGEABLE host buffers (what a numpy caller passes), single process only ----
        t_page = None
        if world == 1:
            t_acc = 0.0
            for _ in range(min(args.steps, 3)):
                if need_flush:
                    flush.zero_()
                    torch.cuda.synchronize()
                t0 = time.perf_counter()
                one_call({k: torch.from_numpy(pts[k]) for k in keys}, bf)
                t_acc += time.perf_counter() - t0
            t_page = t_acc / min(args.steps, 3)
        # small configs finish before nvidia-smi (200 ms period) has reported three times: keep the same workload
        # running, untimed, until it has -- a clocks line without samples verifies nothing
        if rank == 0:
            t_end = time.perf_counter() + 3.0
            while len(sampler.rows) < 4 and time.perf_counter() < t_end:
                one_call(resident, bf)
        clocks = sampler.stop() if rank == 0 else None

    # max over ranks (device-synchronised host clock around synchronous calls)
    if world > 1:
        tt = torch.tensor([t_res, t_e2e, float(np.mean(kern_ms))], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_res, t_e2e, kmean = (float(v) for v in tt.cpu())
        ne = torch.tensor([n_eval], dtype=torch.int64, device=dev)
        dist.all_reduce(ne, op=dist.ReduceOp.SUM)
        n_eval_total = int(ne.item())
    else:
        kmean = float(np.mean(kern_ms))
        n_eval_total = n_eval

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return None

    # ---- workload constants ----
    if stat in ("xi", "wp", "DD", "DDrppi"):
        counts = ref_cell_counts(pts, cfg["L"], st0["nmesh"], dtype)
        n_cand = n_cand_box(counts, st0["refine"])
    elif stat in ("DDsmu", "DDrppi_mocks", "DDsmu_mocks"):
        n_cand = n_cand_data_extent(cfg, pts, N, bins)
    else:
        n_cand = float(n_eval_total)  # no cheap closed form: use the device's own evaluation count
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    sm_max = float(peaks.get("sm_max_mhz", 1965.0))
    lanes = 128 if dtype == np.float32 else 64
    # measured ALU issue rate (tools/ubench.cu: dependent-free FFMA / DFMA streams on every SM, profiles/measured_alu.json)
    # and DRAM traffic of the pair kernel (one ncu capture per config, profiles/measured_traffic.json)
    alu, traffic_tab = {}, {}
    try:
        alu = json.load(open(os.path.join(ROOT, "profiles", "measured_alu.json")))
        traffic_tab = json.load(open(os.path.join(ROOT, "profiles", "measured_traffic.json")))
    except Exception:
        pass
    key = "fp32_warp_instr_per_clk_per_sm" if dtype == np.float32 else "fp64_warp_instr_per_clk_per_sm"
    nominal_wi = lanes / 32.0
    measured_wi = float(alu.get(key, nominal_wi))
    traffic = traffic_tab.get(args.config if not args.npart else "", {}).get("dram_bytes_per_launch")
    flop = FLOP_PER_EVAL[stat]
    # ALU roofline of SURVEY.md 8(d): one evaluation costs INSTR_PER_EVAL lane-instructions of the FP32
    # (FP64) pipe, so peak evals/s = SMs x lanes x clock / INSTR_PER_EVAL; in FLOP terms that mix carries
    # FLOP_PER_EVAL per INSTR_PER_EVAL lane-cycles (sub and mul count 1, fma 2)
    peak_tflops = 148 * measured_wi * 32 * sm_max * 1e6 * flop / INSTR_PER_EVAL[stat] / 1e12
    ngpu_used = max(world, int(lib.cfb_last_device_count()))
    ach_tflops = n_eval_total * flop / (kmean * 1e-3) / 1e12 / ngpu_used  # per GPU
    peak_evals = 148 * measured_wi * 32 * sm_max * 1e6 / INSTR_PER_EVAL[stat]
    line = {
        "metric": "pair evaluations/sec (reference-equivalent candidate pairs, N_cand/t) and DD wall-time",
        "value": n_cand / t_res,
        "unit": "pair_evals/s",
        "n_gpus": ngpu_used,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": t_res * 1e3,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": cfg["dtype"],
        "data": "synthetic",
        "config": workload_config(args, cfg, N, input_bytes),
        "lattice": {"reference_lattice": list(st0["nmesh"]), "refine": list(st0["refine"]), "device_lattice": list(st0["fine"])},
        "timing": {"how": "host clock around synchronous C-ABI calls (device-synchronised on both sides), max over ranks; "
                          "kernel time by CUDA events on the launch stream",
                   "device_ms_per_step": float(np.mean(dev_ms)),
                   "process_model": ("one process, one C call, %d devices inside the library" % ngpu_used) if args.inlib
                   else "one process per GPU"},
        "e2e": {"value": n_cand / t_e2e, "unit": "pair_evals/s", "ms_per_step": t_e2e * 1e3,
                "h2d_bytes_per_step": int(input_bytes), "d2h_bytes_per_step": int(len(bins) * 24 + 64),
                "host_buffers": "pinned", "ms_per_step_pageable": None if t_page is None else t_page * 1e3},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "fp32_alu" if dtype == np.float32 else "fp64_alu", "achieved": ach_tflops, "peak": peak_tflops,
                     "unit": "TFLOP/s", "frac": ach_tflops / peak_tflops, "traffic": traffic,
                     "peak_source": "%s ALU issue rate: 148 SMs x %.2f warp-instr/clk/SM (nominal %.0f) x 32 lanes x %.0f MHz x %d FLOP / %d instr per evaluation (MEASURED_PEAKS.json has no ALU figure; clock = its sm_max_mhz; issue rate from tools/ubench.cu, profiles/measured_alu.json)" % ("measured" if key in alu else "nominal", measured_wi, nominal_wi, sm_max, flop, INSTR_PER_EVAL[stat]),
                     "frac_of_nominal_peak": ach_tflops / (148 * lanes * sm_max * 1e6 * flop / INSTR_PER_EVAL[stat] / 1e12),
                     "kernel_ms": kmean, "gridlink_ms": float(np.mean(grid_ms)), "n_eval": n_eval_total,
                     "evals_per_s": n_eval_total / (kmean * 1e-3), "peak_evals_per_s_per_gpu": peak_evals,
                     "kernel_share_of_step": kmean * 1e-3 / t_res,
                     "levels_per_eval": st0["n_levelpairs"] / max(st0["n_eval"], 1), "n_analytic": st0["n_analytic"],
                     "n_jobs": st0["n_tilepairs"], "n_tiles": st0["n_tiles"], "kernel_kind": st0["kernel_kind"]},
        "n_cand": n_cand,
        "n_pairs_total": tot,
    }
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(cfg, args, budget_s=15.0, n_cand_full=n_cand)
    if world > 1:
        dist.destroy_process_group()
    return line


def ref_call(ref, cfg, pts, n, bins, nthreads):
    """One pass of the unmodified reference on the first n points; returns (seconds, nmesh-free n_cand)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H
    from corrfunc_b200 import _capi

    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    stat = cfg["stat"]
    o = _capi.default_options(dtype, periodic=True, need_avg_sep=bool(cfg.get("avg")), isa=H.ref_isa(),
                              boxsize=cfg["L"] if cfg["L"] > 0 else None)
    w = pts.get("w")
    wt = "pair_product" if cfg.get("weights") else None
    t0 = time.perf_counter()
    if stat == "xi":
        _capi.call_xi(ref, cfg["L"], nthreads, bins, pts["x"][:n], pts["y"][:n], pts["z"][:n], options=o)
    elif stat == "DD":
        _capi.call_DD(ref, 1, nthreads, bins, pts["x"][:n], pts["y"][:n], pts["z"][:n], options=o)
    elif stat == "wp":
        _capi.call_wp(ref, cfg["L"], nthreads, cfg["pimax"], bins, pts["x"][:n], pts["y"][:n], pts["z"][:n], options=o)
    elif stat == "DDrppi":
        _capi.call_DDrppi(ref, 1, nthreads, cfg["pimax"], bins, pts["x"][:n], pts["y"][:n], pts["z"][:n], options=o)
    elif stat == "DDsmu":
        _capi.call_DDsmu(ref, 1, nthreads, bins, cfg["mu_max"], cfg["nmu"], pts["x"][:n], pts["y"][:n], pts["z"][:n],
                         w1=None if w is None else w[:n], weight_type=wt, options=o)
    elif stat == "DDtheta":
        _capi.call_DDtheta(ref, 0, nthreads, bins, pts["ra"][:n], pts["dec"][:n], RA2=pts["ra2"][:n], DEC2=pts["dec2"][:n], options=o)
    elif stat == "DDrppi_mocks":
        o.is_comoving_dist = 1
        _capi.call_DDrppi_mocks(ref, 1, 1, nthreads, cfg["pimax"], bins, pts["ra"][:n], pts["dec"][:n], pts["d"][:n], options=o)
    elif stat == "DDsmu_mocks":
        o.is_comoving_dist = 1
        _capi.call_DDsmu_mocks(ref, 1, 1, nthreads, cfg["mu_max"], cfg["nmu"], bins, pts["ra"][:n], pts["dec"][:n],
                               pts["d"][:n], w1=None if w is None else w[:n], weight_type=wt, options=o)
    return time.perf_counter() - t0


def ref_lattice_for(cfg, n, bins):
    """nmesh/refine the reference picks for a [0,L]^3 periodic box (xi/wp/DD on uniform data)."""
    rmax = bins[-1]
    L = cfg["L"]
    rf = [2, 2, 1]
    if cfg["stat"] == "xi" and rmax < 0.05 * L:
        rf = [1, 1, 1]
    zmax = cfg.get("pimax", rmax)

    def mesh(rf):
        return [max(2, min(100, int(rf[0] * L / rmax))), max(2, min(100, int(rf[1] * L / rmax))),
                max(2, min(100, int(rf[2] * L / zmax)))]

    nm = mesh(rf)
    if (max(nm) <= 10 or n / (nm[0] * nm[1] * nm[2]) >= 250) and max(nm) < 100:
        rf = [rf[0] + 1, rf[1] + 1, rf[2]]
        nm = mesh(rf)
    return nm, rf


def cpu_baseline(cfg, args, budget_s=15.0, steps=1, n_cand_full=None):
    """The reference's OpenMP AVX-512 path on this host's cores, on a bounded subsample."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H

    ref = H.load_ref()
    if ref is None:
        return {"value": None, "unit": "pair_evals/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref not built"}
    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    cores = os.cpu_count() or 1
    bins = make_bins(cfg["bins"])
    stat = cfg["stat"]
    n_full = args.npart or cfg["N"]
    # calibrate on a small subsample, then size the sample for ~budget_s (cost ~ n^2)
    n0 = min(n_full, 1_000_000 if stat != "DDtheta" else 300_000)
    pts = gen_points(cfg, min(n_full, 12_000_000), dtype)
    t_cal = ref_call(ref, cfg, pts, n0, bins, cores)
    n_s = int(min(len(next(iter(pts.values()))), n0 * max(1.0, (budget_s / max(t_cal, 1e-3)) ** 0.5)))
    ts = [ref_call(ref, cfg, pts, n_s, bins, cores) for _ in range(steps)]
    t = float(np.mean(ts))
    if stat in ("xi", "wp", "DD"):
        nm, rf = ref_lattice_for(cfg, n_s, bins)
        counts = ref_cell_counts({k: pts[k][:n_s] for k in "xyz"}, cfg["L"], nm, dtype)
        n_cand = n_cand_box(counts, rf)
    elif stat in ("DDsmu", "DDrppi_mocks", "DDsmu_mocks"):
        n_cand = n_cand_data_extent(cfg, pts, n_s, bins)
    elif n_cand_full:
        # no closed form for this lattice: the device's own count of the full workload, scaled to the sample
        # (candidate pairs of a uniform catalogue grow with the square of the point count)
        n_cand = float(n_cand_full) * (n_s / float(n_full)) ** 2
    else:
        n_cand = float("nan")
    return {"value": n_cand / t, "unit": "pair_evals/s", "cores": cores, "kind": "reference",
            "isa": "avx512f" if H.ref_variant() == "v4" else "avx", "seconds": t, "n_cand": n_cand,
            "omp": {"proc_bind": os.environ.get("OMP_PROC_BIND"), "places": os.environ.get("OMP_PLACES")},
            "sample": "first %d of the %d points (same box, same bins): reference %s, %d OpenMP threads" % (n_s, n_full, stat, cores)}


def n_cand_of(cfg, pts, n, bins):
    """Candidate pairs of the reference's cell-pair set for the first n points (None when there is no cheap closed form)."""
    stat = cfg["stat"]
    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    if stat in ("xi", "wp", "DD", "DDrppi"):
        nm, rf = ref_lattice_for(cfg, n, bins)
        return n_cand_box(ref_cell_counts({k: pts[k][:n] for k in "xyz"}, cfg["L"], nm, dtype), rf)
    if stat in ("DDsmu", "DDrppi_mocks", "DDsmu_mocks"):
        return n_cand_data_extent(cfg, pts, n, bins)
    return None


def run_reference(args, cfg):
    """The reference arm: the UNMODIFIED reference (oracle/_ref, AVX-512F kernels, every host thread) on the workload of
    the GPU arm.  When K full-size steps fit the time budget they are run as they are.  Otherwise (config 5: 100 M points
    take the reference 10-25 minutes per step) the reference is timed at two sizes of the same box and bins -- K steps at
    n_small, one at n_large (10 M and 30 M when the budget allows, the sizes SURVEY 8(d) names) -- the exponent of
    t ~ n^p is fitted from the two, and the full-size time is extrapolated: `ms_per_step` and `value` are those of the
    FULL workload and carry `extrapolated: true`, the measured seconds and the fit sit beside them."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import harness as H

    ref = H.load_ref()
    N = args.npart or cfg["N"]
    base = {"impl": "reference", "metric": "pair evaluations/sec (reference-equivalent candidate pairs, N_cand/t) and DD wall-time",
            "unit": "pair_evals/s", "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": workload_config(args, cfg, N, input_bytes_of(cfg, N))}
    if ref is None:
        base.update(value=None, ms_per_step=None, cpu_baseline={"value": None, "kind": "reference", "cores": 0,
                                                                "sample": "oracle/_ref not built"})
        return base
    dtype = np.float32 if cfg["dtype"] == "f32" else np.float64
    cores = os.cpu_count() or 1
    bins = make_bins(cfg["bins"])
    stat = cfg["stat"]
    K = max(1, args.steps)
    budget = float(os.environ.get("CORRFUNC_BENCH_REF_BUDGET_S", "200"))
    pts = gen_points(cfg, N, dtype)
    n_cal = min(N, 1_000_000 if stat != "DDtheta" else 300_000)
    t_cal = ref_call(ref, cfg, pts, n_cal, bins, cores)
    t_full_guess = t_cal * (N / n_cal) ** 2
    n_cand_full = n_cand_of(cfg, pts, N, bins)
    isa = "avx512f" if H.ref_variant() == "v4" else "avx"
    if (K + min(args.warmup, 1)) * t_full_guess <= budget:
        for _ in range(min(args.warmup, 1)):
            ref_call(ref, cfg, pts, N, bins, cores)
        ts = [ref_call(ref, cfg, pts, N, bins, cores) for _ in range(K)]
        t = float(np.mean(ts))
        extra = {"extrapolated": False, "seconds_per_step": ts}
        sample = "the full workload: %d points, reference %s, %d OpenMP threads, %d steps" % (N, stat, cores, K)
    else:
        # two sizes: K steps at n_small (45 % of the budget), one at n_large (45 %)
        n_small = int(min(10_000_000, n_cal * (0.45 * budget / (K * t_cal)) ** 0.5, N))
        n_large = int(min(30_000_000, n_cal * (0.45 * budget / t_cal) ** 0.5, N))
        if n_large < 1.5 * n_small:
            n_small = int(n_large / 1.5)
        ts = [ref_call(ref, cfg, pts, n_small, bins, cores) for _ in range(K)]
        t_small = float(np.mean(ts))
        t_large = ref_call(ref, cfg, pts, n_large, bins, cores)
        p_fit = float(np.log(t_large / t_small) / np.log(n_large / n_small))
        t = t_large * (N / n_large) ** p_fit
        # error bar: the exponent is a two-point fit; +-0.05 on it (the run-to-run spread of the two timings) moves the
        # extrapolation by the factor (N / n_large)^0.05
        err = (N / n_large) ** 0.05 - 1.0
        extra = {"extrapolated": True, "fit": {"n": [n_small, n_large], "seconds": [t_small, t_large], "exponent": p_fit,
                                                "steps_at_n_small": K, "relative_error_of_extrapolation": err,
                                                "model": "t = t(n_large) * (N / n_large)^exponent, same box and bins at lower density"}}
        sample = ("first %d (x%d steps) and first %d (x1) of the %d points, same box and bins; t ~ n^%.2f; full size extrapolated "
                  "(+-%.0f %%): reference %s, %d OpenMP threads" % (n_small, K, n_large, N, p_fit, 100 * err, stat, cores))
    value = None if n_cand_full is None else n_cand_full / t
    if value is None:
        # no closed form for this lattice (DDtheta): pairs in range per second over the whole catalogue are not comparable
        # with the GPU arm's unit; report wall time and leave the rate to the GPU arm's own count
        value = float("nan")
    cb = {"value": value, "unit": "pair_evals/s", "cores": cores, "kind": "reference", "isa": isa, "seconds": t,
          "n_cand": n_cand_full, "omp": {"proc_bind": os.environ.get("OMP_PROC_BIND"), "places": os.environ.get("OMP_PLACES")},
          "sample": sample}
    cb.update(extra)
    base.update(value=value, ms_per_step=t * 1e3, cpu_baseline=cb,
                e2e={"value": value, "unit": "pair_evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})
    base.update({k: v for k, v in extra.items() if k in ("extrapolated", "fit")})
    return base


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default=os.environ.get("CORRFUNC_BENCH_CONFIG", "c5"))
    ap.add_argument("--npart", type=int, default=0, help="override N (marks the line as reduced)")
    ap.add_argument("--same-density", action="store_true", help="with --npart: shrink the box to keep the number density")
    ap.add_argument("--occ", type=int, default=0, help="target particles per device cell (0 = default)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--inlib", action="store_true",
                    help="single process: let the library shard one call over --gpus devices (CORRFUNC_B200_NGPUS)")
    args = ap.parse_args()
    if args.impl == "reference":
        # the reference's OpenMP threads pinned to cores.  Only this arm: it is a process of its own
        # (rank 0), and libgomp reads these when it is first loaded -- in the GPU arm, whose ranks share the host,
        # pinning every rank's initial thread to the first core would serialise them.
        os.environ.setdefault("OMP_PROC_BIND", "close")
        os.environ.setdefault("OMP_PLACES", "cores")
    cfg = dict(CONFIGS[args.config])
    if args.npart and args.same_density and cfg["L"] > 0:
        cfg["L"] = float(cfg["L"] * (args.npart / cfg["N"]) ** (1.0 / 3.0))
    # stdout carries exactly one JSON line: anything a library prints there meanwhile (e.g. NCCL's version
    # banner) is routed to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = run_reference(args, cfg) if args.impl == "reference" else run_ours(args, cfg)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        os.close(real_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
