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


def channel_tones(c, snap=None):
    f1, f2 = 400.0 + 7.0 * (c % 97), 1500.0 + 11.0 * (c % 89)
    if snap:
        f1, f2 = np.rint(f1 / snap) * snap, np.rint(f2 / snap) * snap
    return f1, f2


def baseband(mode, c, t, snap=None):
    """Complex baseband modulation s_c(t) for channel index c (snap: tone frequencies rounded to multiples of it)."""
    f1, f2 = channel_tones(c, snap)
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
        if snap:
            fo = np.rint(fo / snap) * snap
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


def syn_iq_fft(fs, n, modes, carriers, seed, decim, noise_db=-40.0, total_amp=16000.0, active=None, impulses=0):
    """SYN-IQ for the full-size banks (1024-4096 carriers x 20-40 M samples), built in the frequency domain:
    every channel's modulation s_c is evaluated at the low rate fs/decim (n/decim samples), transformed, and its
    spectrum placed at the carrier's bin of ONE n-point inverse FFT -- the same x[n] = sum_c A s_c e^{j(2 pi f_c n/fs + p_c)}
    + noise as syn_iq, with s_c band-limited-interpolated instead of evaluated per wideband sample (minutes -> seconds).
    The stream is exactly periodic in n samples: carriers and modulation tones are snapped to multiples of fs/n, and the
    snapped carriers are returned (tune the receivers to those). Every carrier gets a seeded random phase p_c, so a
    thousand carriers do not add up coherently at n = 0 and the sum stays inside the int16 range.
    `active` limits the carriers that are actually present (default: all); `impulses` single-sample spikes of
    amplitude 30000 are added at seeded positions (noise-blanker stimulus). decim must divide n, and fs/decim must
    cover the widest modulation (FM: ~ +-5 kHz). Returns (complex64[n], carriers)."""
    from scipy import fft as sfft
    nch = len(modes)
    assert n % decim == 0
    m = n // decim
    A = total_amp / np.sqrt(nch)
    sigma = A * 10.0 ** (noise_db / 20.0)
    df = fs / n
    kbin = np.rint(np.asarray(carriers, dtype=np.float64) / df).astype(np.int64)
    snapped = kbin * df
    rng = np.random.Generator(np.random.PCG64([seed, n]))
    phase = rng.uniform(0.0, 2.0 * np.pi, nch)
    X = np.zeros(n, dtype=np.complex128)
    t = np.arange(m, dtype=np.float64) * (decim / fs)
    q = np.fft.fftfreq(m, 1.0 / m).astype(np.int64)          # baseband bin numbers 0..m/2-1, -m/2..-1
    for c in (range(nch) if active is None else active):
        b = baseband(modes[c], c, t, snap=df) * np.exp(1j * phase[c])
        X[(kbin[c] + q) % n] += sfft.fft(b) * (A * n / m)
    x = sfft.ifft(X, workers=-1, overwrite_x=True)
    del X
    out = np.empty(n, dtype=np.complex64)
    step = 1 << 22
    for s0 in range(0, n, step):
        k = min(step, n - s0)
        g = rng.standard_normal((k, 2), dtype=np.float32)
        out[s0:s0 + k] = x[s0:s0 + k].astype(np.complex64)
        out[s0:s0 + k] += np.float32(sigma) * g.view(np.complex64)[:, 0]
    if impulses:
        pos = np.sort(rng.choice(n, size=impulses, replace=False))
        out[pos] += np.complex64(30000.0)
    return out, snapped


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
