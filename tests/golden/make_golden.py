#!/usr/bin/env python3
"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libcutesdr_ref*.so, built by
oracle/Makefile from /root/reference). Run in the build container only; the fixtures are committed so
the GPU box (which has no /root/reference) can check both the oracle restatement and the CUDA path
against outputs of the reference itself.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from cutesdr_b200 import modes as M  # noqa: E402
from cutesdr_b200.synth import carrier_grid, syn_iq  # noqa: E402
from oracle import ref_binding as rb  # noqa: E402


def golden_inputs():
    """Deterministic inputs shared by the generator and the tests (never stored: regenerated)."""
    fs = 2e6
    n = 300000
    t = np.arange(n) / fs
    s = 1 + 0.5 * np.cos(2 * np.pi * 1000 * t) + 0.3 * np.cos(2 * np.pi * 1700 * t)
    am = (8000 * s * np.exp(2j * np.pi * 250000 * t)).astype(np.complex64)
    modes = [M.DEMOD_AM, M.DEMOD_SAM, M.DEMOD_FM, M.DEMOD_USB, M.DEMOD_LSB, M.DEMOD_CWU]
    carriers = carrier_grid(len(modes), 150e3)
    mixed = syn_iq(fs, 700000, modes, carriers, seed=20260)
    rng = np.random.default_rng(424242)
    noise = (3000 * (rng.standard_normal(8192) + 1j * rng.standard_normal(8192))).astype(np.complex64)
    return dict(fs=fs, am=am, modes=modes, carriers=carriers, mixed=mixed, noise=noise)


def info_for(mode):
    if mode == M.DEMOD_USB:
        return M.demod_info(mode, HiCut=2800, LowCut=100)
    if mode == M.DEMOD_LSB:
        return M.demod_info(mode, HiCut=-100, LowCut=-2800)
    if mode == M.DEMOD_CWU:
        return M.demod_info(mode, HiCut=250, LowCut=-250, Offset=700)
    return M.demod_info(mode)


def main():
    g = golden_inputs()
    fs = g["fs"]
    out = {}
    # stage ladders + block lengths
    ladders = []
    for rate in (2e6, 20e6, 100147200.0, 200294400.0):
        for bw in (10000, 15000, 20000):
            d = rb.RefDownConvert(big=True)
            r = d.SetDataRate(rate, bw)
            ladders.append((rate, bw, r, d.stages()))
    out["ladder_rate"] = np.array([l[0] for l in ladders])
    out["ladder_bw"] = np.array([l[1] for l in ladders])
    out["ladder_out"] = np.array([l[2] for l in ladders])
    out["ladder_stages"] = np.array([",".join(map(str, l[3])) for l in ladders])
    # config 1: single-channel AM on the two-tone stream, then 48 kHz
    d = rb.RefDemodulator()
    d.SetInputSampleRate(fs)
    d.SetDemod(M.DEMOD_AM, M.demod_info(M.DEMOD_AM))
    d.SetDemodFreq(-250000)
    y, taps = d.run(g["am"], taps=(1, 2, 3))
    out["am_audio"] = y
    out["am_tap1"] = taps[1][:2 * 2048]
    out["am_tap2"] = taps[2][:2 * 2048]
    out["am_tap3"] = taps[3][:2 * 2048]
    out["am_smeter_ave"] = np.array([d.GetSMeterAve()])
    r = rb.RefFractResampler(8192)
    out["am_audio48"] = np.concatenate([r.Resample(y[k:k + 1024], 31250.0 / 48000.0) for k in range(0, len(y), 1024)])
    # every mode on the mixed stream
    for c, mode in enumerate(g["modes"]):
        d = rb.RefDemodulator()
        d.SetInputSampleRate(fs)
        d.SetDemod(mode, info_for(mode))
        d.SetDemodFreq(-g["carriers"][c])
        out["mixed_audio_%d" % c] = d.run(g["mixed"])
    # display FFT: 4 frames of the mixed stream, ave 2
    f = rb.RefFft()
    f.SetFFTParams(4096, False, 0.0, fs)
    f.SetFFTAve(2)
    screens = []
    for k in range(4):
        f.PutInDisplayFFT(g["mixed"][k * 4096:(k + 1) * 4096])
        ov, s1 = f.GetScreenIntegerFFTData(255, 800, 0.0, -140.0, -1000000, 1000000)
        ov2, s2 = f.GetScreenIntegerFFTData(600, 1000, 0.0, -140.0, -50000, 50000)
        screens.append(np.concatenate([s1, s2, [int(ov), int(ov2)]]))
    out["fft_screens"] = np.array(screens, dtype=np.int32)
    out["fft_ave"] = f.avebuf()
    # stand-alone stages on the noise vector
    a = rb.RefAgc()
    a.SetParameters(1, 0, -100, 30, 0, 200, 48900.0)
    out["agc_out"] = np.concatenate([a.ProcessData(g["noise"][k:k + 1024].astype(np.complex128)) for k in range(0, 4096, 1024)])
    ff = rb.RefFastFIR()
    ff.SetupParameters(100, 2800, 0, 62500.0)
    out["fastfir_out"] = ff.ProcessData(g["noise"].astype(np.complex128))
    nb = rb.RefNoiseProc()
    nb.SetupBlanker(True, 50.0, 50.0, fs)
    x = g["noise"].astype(np.complex128).copy()
    x[[1000, 5000]] += 30000.0
    x = x.astype(np.complex64).astype(np.complex128)
    out["blanker_out"] = nb.ProcessBlanker(x)
    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    print("wrote", os.path.join(HERE, "reference_outputs.npz"), os.path.getsize(os.path.join(HERE, "reference_outputs.npz")), "bytes")


if __name__ == "__main__":
    main()
