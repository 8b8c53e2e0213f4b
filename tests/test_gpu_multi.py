"""Full-size-density parity at more than one GPU (skipped on a single-GPU box): config 5 at the same density with 10 M
points against the committed output of the unmodified reference (tests/golden/ref_fullsize_c5sd10M.npz), through both
multi-GPU process models:
  * one C call sharded inside the library over every visible device (CORRFUNC_B200_NGPUS);
  * one process per GPU under torchrun, histograms all-reduced over NCCL (corrfunc_b200.parallel)."""
import os
import subprocess
import sys

import numpy as np
import pytest

import harness as H

pytestmark = pytest.mark.gpu

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import bench
from corrfunc_b200 import _capi, _lib, parallel
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["CORRFUNC_B200_DEVICE"] = str(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
parallel.enable_distributed(dist, dev)
cfg = bench.config_by_name("c5sd10M")
pts = bench.gen_points(cfg, cfg["N"], np.float32)
o = _capi.default_options(np.float32, periodic=True, boxsize=cfg["L"])
r = _capi.call_xi(_lib.load(), cfg["L"], 1, bench.make_bins(cfg["bins"]), pts["x"], pts["y"], pts["z"], options=o)
want = np.load(os.path.join(sys.argv[1], "tests", "golden", "ref_fullsize_c5sd10M.npz"))["npairs"].astype(np.uint64)
ok = np.array_equal(np.asarray(r["npairs"], dtype=np.uint64), want)
print("RANK %d OK=%d" % (dist.get_rank(), int(ok)), flush=True)
dist.destroy_process_group()
sys.exit(0 if ok else 3)
'''


def _ndev():
    import torch

    return torch.cuda.device_count()


def test_c5sd10M_golden_one_call_all_gpus_inside_the_library():
    if _ndev() < 2:
        pytest.skip("needs at least 2 GPUs")
    import bench
    from corrfunc_b200 import _capi, _lib

    cfg = bench.config_by_name("c5sd10M")
    pts = bench.gen_points(cfg, cfg["N"], np.float32)
    o = _capi.default_options(np.float32, periodic=True, boxsize=cfg["L"])
    lib = _lib.load()
    saved = {k: os.environ.pop(k, None) for k in ("CORRFUNC_B200_NGPUS", "CORRFUNC_B200_DEVICE")}
    try:
        os.environ["CORRFUNC_B200_NGPUS"] = str(_ndev())
        r = _capi.call_xi(lib, cfg["L"], 1, bench.make_bins(cfg["bins"]), pts["x"], pts["y"], pts["z"], options=o)
        assert lib.cfb_last_device_count() == _ndev()
    finally:
        os.environ.pop("CORRFUNC_B200_NGPUS", None)
        for k, v in saved.items():
            if v is not None:
                os.environ[k] = v
    want = np.load(os.path.join(H.GOLDEN, "ref_fullsize_c5sd10M.npz"))["npairs"].astype(np.uint64)
    assert np.array_equal(np.asarray(r["npairs"], dtype=np.uint64), want)


def test_c5sd10M_golden_one_process_per_gpu_torchrun(tmp_path):
    n = _ndev()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ)
    env.pop("CORRFUNC_B200_NGPUS", None)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n),
                        "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), H.ROOT],
                       capture_output=True, text=True, env=env, timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-2000:]
    assert p.stdout.count("OK=1") == n


@pytest.mark.gpu
def test_interrupt_stops_a_running_count(tmp_path):
    """SIGINT during a long count (utils/macros.h:145-167, theory/DD/countpairs_impl.c.src:475-477,554-569): the library's
    handler prints the reference's message and sets the mapped flag, the persistent warps stop fetching tiles, the call
    returns EXIT_FAILURE (RuntimeError in the wrapper) well before the count would have finished."""
    import signal
    import time

    script = tmp_path / "long_count.py"
    script.write_text(
        "import os, sys, time\n"
        "import numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from corrfunc_b200.theory import xi\n"
        "N, L = 40_000_000, 1473.0\n"  # config 5's density: about 2 s of pair counting
        "rng = np.random.default_rng(5)\n"
        "x, y, z = (rng.random(N, dtype=np.float32) * np.float32(L) for _ in range(3))\n"
        "bins = np.logspace(-1, np.log10(150.0), 31)\n"
        "xi(L, 1, bins, x[:100000], y[:100000], z[:100000])\n"  # context, kernels loaded
        "print('START', flush=True)\n"
        "t0 = time.time()\n"
        "try:\n"
        "    xi(L, 1, bins, x, y, z)\n"
        "    print('FINISHED %%.3f' %% (time.time() - t0), flush=True)\n"
        "except RuntimeError as e:\n"
        "    print('INTERRUPTED %%.3f' %% (time.time() - t0), flush=True)\n" % H.ROOT)
    env = dict(os.environ, CORRFUNC_B200_NGPUS="1")
    p = subprocess.Popen([sys.executable, str(script)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
    line = p.stdout.readline()
    assert line.startswith("START"), line + p.stderr.read()
    time.sleep(0.8)  # copies, upload and gridlink take 0.2-0.4 s; the pair kernel (about 2 s) is running now
    p.send_signal(signal.SIGINT)
    t_sig = time.time()
    out, err = p.communicate(timeout=120)
    waited = time.time() - t_sig
    assert "INTERRUPTED" in out, out + err
    assert "Received signal" in err and "Aborting" in err
    assert waited < 1.0, "the count did not stop early: %.2f s after the signal" % waited
