"""The multi-GPU side of the library (pumas_b200/csrc/comm.cu): the C example
examples/allreduce_ranks.c drives one rank per GPU from a single C process
(one host thread each), accumulates the tallies of two batches per rank on the
device and sums them over the ranks with ONE in-library NCCL all-reduce.

CPU: the example compiles and links against include/*.h + libpumas_b200.so.
GPU: it runs with as many ranks as there are GPUs (1 on the driver's GPU tier,
where NCCL still initialises a communicator; 2+ under `gpurun --gpus N`), and
the segment tally / accumulation kernels are checked against numpy."""
import os
import shutil
import subprocess

import numpy as np
import pytest

import harness
import pumas_b200 as pb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "pumas_b200")


def _cuda_home():
    for d in (os.environ.get("CUDA_HOME"), "/usr/local/cuda"):
        if d and os.path.exists(os.path.join(d, "include",
                                             "cuda_runtime_api.h")):
            return d
    return None


def _build(tmp_path):
    cuda = _cuda_home()
    if shutil.which("gcc") is None or cuda is None:
        pytest.skip("no gcc / CUDA toolkit")
    exe = str(tmp_path / "allreduce_ranks")
    subprocess.run(
        ["gcc", "-std=c99", "-O2", "-Wall", "-Wextra", "-Werror",
         os.path.join(ROOT, "examples", "allreduce_ranks.c"),
         "-I" + os.path.join(ROOT, "include"),
         "-I" + os.path.join(cuda, "include"), "-L" + LIBDIR,
         "-L" + os.path.join(cuda, "lib64"), "-lpumas_b200", "-lcudart",
         "-lpthread", "-Wl,-rpath," + LIBDIR,
         "-Wl,-rpath," + os.path.join(cuda, "lib64"), "-o", exe], check=True)
    return exe


def test_example_compiles_and_links(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 2 and "usage" in r.stderr


@pytest.mark.gpu
def test_ranks_reduce_in_library(tmp_path):
    import torch
    exe = _build(tmp_path)
    ranks = min(torch.cuda.device_count(), 8)
    env = dict(os.environ)
    # the NCCL the process would load: the one bundled with PyTorch
    import glob
    import site
    for sp in site.getsitepackages():
        hits = glob.glob(os.path.join(sp, "nvidia", "nccl", "lib",
                                      "libnccl.so.2"))
        if hits:
            env["PUMAS_B200_NCCL"] = hits[0]
    r = subprocess.run([exe, harness.dump_path("materials.muon"), str(ranks)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert ("ranks %d: tracks %d" % (ranks, ranks * 2 * 20000)) in r.stdout


@pytest.mark.gpu
def test_accumulate_and_segments():
    import torch
    n_bins, per_bin = 37, 50
    n = n_bins * per_bin
    rng = np.random.default_rng(3)
    states = pb.make_states(n, 1.)
    states["weight"] = rng.uniform(0.5, 2., n)
    events_np = rng.integers(0, 2, n).astype(np.int32) * pb.EVENT_LIMIT_DISTANCE
    dev = torch.from_numpy(states.view(np.uint8).reshape(n, 104).copy()).cuda()
    events = torch.from_numpy(events_np).cuda()
    out = torch.zeros(3 * n_bins, dtype=torch.float64, device="cuda")
    pb.tally_segments(dev.data_ptr(), events.data_ptr(), n_bins, per_bin,
                      pb.EVENT_LIMIT_DISTANCE, out.data_ptr())
    total = torch.ones(3 * n_bins, dtype=torch.float64, device="cuda")
    pb.tally_accumulate(total.data_ptr(), out.data_ptr(), 3 * n_bins)
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(n_bins, 3)
    w = states["weight"].reshape(n_bins, per_bin)
    sel = (events_np != 0).reshape(n_bins, per_bin)
    assert np.allclose(got[:, 0], (w * sel).sum(axis=1), rtol=1e-14)
    assert np.allclose(got[:, 1], (w * w * sel).sum(axis=1), rtol=1e-14)
    assert np.array_equal(got[:, 2], sel.sum(axis=1))
    assert np.array_equal(total.cpu().numpy(), 1. + out.cpu().numpy())
    # twice the same input: bit identical (fixed reduction order)
    out2 = torch.zeros_like(out)
    pb.tally_segments(dev.data_ptr(), events.data_ptr(), n_bins, per_bin,
                      pb.EVENT_LIMIT_DISTANCE, out2.data_ptr())
    torch.cuda.synchronize()
    assert torch.equal(out, out2)
