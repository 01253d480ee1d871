"""Synthetic wideband IQ ("SYN-IQ") used by tests and bench.py.

Generated in float64 and rounded to complex64: the GPU consumes the complex64
values, the CPU oracle consumes the same values widened to double. Full scale
is int16 (+-32767), as every dB constant in the reference assumes.
"""
import numpy as np

from .modes import DEMOD_AM, DEMOD_SAM, DEMOD_FM, DEMOD_USB, DEMOD_LSB, DEMOD_CWU, DEMOD_CWL


def carrier_grid(nch, spacing):
    """f_c = (c - Nch/2 + 1/2) * spacing; a channel at +f_c is tuned with SetDemodFreq(-f_c)."""
    return (np.arange(nch) - nch / 2 + 0.5) * spacing


def channel_tones(c):
    return 400.0 + 7.0 * (c % 97), 1500.0 + 11.0 * (c % 89)


def baseband(mode, c, t):
    """Complex baseband modulation s_c(t) for channel index c."""
    f1, f2 = channel_tones(c)
    if mode in (DEMOD_AM, DEMOD_SAM):
        return (1.0 + 0.5 * np.cos(2 * np.pi * f1 * t) + 0.3 * np.cos(2 * np.pi * f2 * t)).astype(np.complex128)
    if mode == DEMOD_FM:
        # two-tone FM, peak deviation 2.5 kHz split 60/40 between the tones
        b1 = 0.6 * 2500.0 / f1
        b2 = 0.4 * 2500.0 / f2
        return np.exp(1j * (b1 * np.sin(2 * np.pi * f1 * t) + b2 * np.sin(2 * np.pi * f2 * t)))
    if mode in (DEMOD_CWU, DEMOD_CWL):
        # carrier plus a weak +-(60 + c mod 50) Hz sideband: both inside a +-250 Hz CW filter
        fo = 60.0 + (c % 50)
        return 1.0 + 0.2 * np.exp(2j * np.pi * fo * t)
    sign = 1.0 if mode == DEMOD_USB else -1.0
    return 0.5 * np.exp(sign * 2j * np.pi * f1 * t) + 0.5 * np.exp(sign * 2j * np.pi * f2 * t)


def syn_iq(fs, n, modes, carriers, seed, n0=0, noise_db=-40.0, total_amp=16000.0, chunk=1 << 20):
    """x[n] = sum_c A s_c(n/fs) e^{j 2 pi f_c n/fs} + sigma (g_I + j g_Q),  A = total_amp/sqrt(Nch),
    sigma = A*10^(noise_db/20), Gaussian g from PCG64(seed). Returns complex64[n] starting at
    absolute sample index n0 (so a stream can be generated block by block)."""
    nch = len(modes)
    A = total_amp / np.sqrt(nch)
    sigma = A * 10.0 ** (noise_db / 20.0)
    out = np.empty(n, dtype=np.complex64)
    rng = np.random.Generator(np.random.PCG64([seed, n0]))
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        idx = np.arange(n0 + s, n0 + s + m, dtype=np.float64)
        t = idx / fs
        acc = np.zeros(m, dtype=np.complex128)
        for c in range(nch):
            # phase in turns reduced mod 1 before the exp keeps float64 accuracy at long offsets
            ph = np.mod(carriers[c] / fs * idx, 1.0)
            acc += baseband(modes[c], c, t) * np.exp(2j * np.pi * ph)
        acc *= A
        acc += sigma * (rng.standard_normal(m) + 1j * rng.standard_normal(m))
        out[s:s + m] = acc.astype(np.complex64)
    return out


def snr_db(ref, test):
    """10 log10( sum|ref|^2 / sum|ref-test|^2 )."""
    ref = np.asarray(ref)
    test = np.asarray(test)
    num = float(np.sum(np.abs(ref) ** 2))
    den = float(np.sum(np.abs(ref - test) ** 2))
    if den == 0.0:
        return float("inf")
    if num == 0.0:
        return float("-inf")
    return 10.0 * np.log10(num / den)
