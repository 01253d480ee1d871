"""Drop-in check through the C++ shims: tests/cpp/compat_receiver.cpp uses only the reference's class API.
The same source is built against cutesdr_b200/compat (GPU) and against the reference's own dsp/ sources
(oracle/_ref/ref_receiver, CPU); their outputs must agree."""
import os
import struct
import subprocess

import numpy as np
import pytest

from cutesdr_b200 import modes as M
from cutesdr_b200.synth import snr_db, syn_iq

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU_BIN = os.path.join(ROOT, "tests", "cpp", "compat_receiver")
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "ref_receiver")


def read_out(path):
    b = open(path, "rb").read()
    na = struct.unpack_from("i", b, 0)[0]
    audio = np.frombuffer(b, dtype=np.float64, count=na, offset=4)
    off = 4 + 8 * na
    n48 = struct.unpack_from("i", b, off)[0]
    a48 = np.frombuffer(b, dtype=np.int16, count=n48, offset=off + 4)
    off += 4 + 2 * n48
    w = struct.unpack_from("i", b, off)[0]
    screen = np.frombuffer(b, dtype=np.int32, count=w, offset=off + 4)
    ov = struct.unpack_from("i", b, off + 4 + 4 * w)[0]
    return audio, a48, screen, ov


@pytest.mark.parametrize("mode,lo,hi", [(M.DEMOD_AM, -5000, 5000), (M.DEMOD_USB, 100, 2800)])
def test_same_host_source_against_shims_and_reference(tmp_path, mode, lo, hi):
    if not os.path.exists(GPU_BIN) or not os.path.exists(REF_BIN):
        pytest.skip("tests/cpp binaries not built (run __graft_entry__.build())")
    fs, fc = 2e6, 250000.0
    iq = syn_iq(fs, 400000, [mode], [fc], seed=20261, total_amp=8000.0)
    p = tmp_path / "iq.c64"
    iq.tofile(p)
    outs = []
    for exe, name in ((REF_BIN, "ref.bin"), (GPU_BIN, "gpu.bin")):
        o = tmp_path / name
        subprocess.check_call([exe, str(p), str(fs), str(mode), str(lo), str(hi), str(-fc), "4096", str(o)])
        outs.append(read_out(o))
    (ra, r48, rs, rov), (ga, g48, gs, gov) = outs
    assert len(ra) == len(ga) > 4096
    assert snr_db(ra, ga) > 90.0
    assert len(r48) == len(g48) and np.max(np.abs(r48.astype(np.int32) - g48.astype(np.int32))) <= 1
    assert rov == gov and np.max(np.abs(rs - gs)) <= 1


def test_cpp_multi_gpu_host_runs_through_the_c_abi(tmp_path):
    """tests/cpp/mgpu_host.cpp (C ABI only: cutesdr_mgpu_*, cutesdr_bank_process_async_bcast, cutesdr_host_alloc) on the
    GPUs of this box: one process per GPU when there are two, a world of one otherwise."""
    exe = os.path.join(ROOT, "tests", "cpp", "mgpu_host")
    if not os.path.exists(exe):
        pytest.skip("tests/cpp binaries not built (run __graft_entry__.build())")
    import torch
    world = 2 if torch.cuda.device_count() >= 2 else 1
    idf = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([exe, str(r), str(world), idf, "256", "8"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in range(world)]
    outs = [p.communicate(timeout=300) for p in procs]
    for r, (p, (out, err)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, (r, out, err)
        assert "rank %d/%d" % (r, world) in out and "8 blocks" in out
        # 8 blocks of ~10 ms at 48 kHz: three or four FIR bursts of ~1002 resampled samples each have completed
        n = int(out.strip().split(",")[-1].split()[0])
        assert 2500 < n < 4500, out
