"""Debug probe: move one channel between chains on a running bank and compare with a fresh oracle receiver."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import cutesdr_b200 as cs
from cutesdr_b200 import modes as M
from cutesdr_b200.synth import carrier_grid, snr_db
from oracle import oracle_binding as ob
from tests.test_gpu_parity import _cheap_wideband

fs, nch, nblk = 2e6, 8, 60
base = [[M.DEMOD_AM, M.DEMOD_USB, M.DEMOD_FM, M.DEMOD_CWU][c % 4] for c in range(nch)]
carriers = carrier_grid(nch, 11000.0)

def info(m, agc=True):
    d = {M.DEMOD_AM: (-5000, 5000), M.DEMOD_USB: (100, 2800), M.DEMOD_FM: (-5000, 5000), M.DEMOD_CWU: (-250, 250)}[m]
    return M.demod_info(m, LowCut=d[0], HiCut=d[1], Offset=700 if m == M.DEMOD_CWU else 0, AgcOn=agc, AgcManualGain=60)

for seq in ([M.DEMOD_USB], [M.DEMOD_CWU], [M.DEMOD_USB, M.DEMOD_CWU], [M.DEMOD_USB, M.DEMOD_CWU, M.DEMOD_AM]):
    b = cs.ReceiverBank(nch, fs)
    for c in range(nch):
        b.SetDemod(c, base[c], info(base[c]))
        b.SetDemodFreq(c, -carriers[c])
    L = b.block_length()
    iq = _cheap_wideband(L, nblk, seed=77)
    got = []
    k0 = None
    for k in range(nblk):
        if k >= 5 and k - 5 < len(seq) * 3 and (k - 5) % 3 == 0:
            m = seq[(k - 5) // 3]
            i2 = info(m, agc=False)
            b.SetDemod(0, m, i2)
            b.SetDemodFreq(0, -carriers[0])
            got = []
            k0 = k
        a, n = b.ProcessData(iq[k * L:(k + 1) * L])
        got.append(a[0, :n[0]].copy())
    d = ob.Demodulator()
    d.SetInputSampleRate(fs)
    d.SetDemod(m, i2)
    d.SetDemodFreq(-carriers[0])
    exp = d.run(iq[k0 * L:])
    y = np.concatenate(got)
    print("seq", [M.MODE_NAMES[x] for x in seq], "k0", k0, "len", len(exp), len(y), "rate", b.GetOutputRate(0))
    best = (-1e9, 0)
    w0 = 3072 if len(exp) > 7000 else 1100
    for lag in range(-1100, 1101):
        a = exp[w0:w0 + 1024]
        bb = y[w0 + lag:w0 + lag + 1024]
        if len(bb) == len(a) and w0 + lag >= 0:
            s = snr_db(a, bb)
            if s > best[0]:
                best = (s, lag)
    print("   best snr %.1f dB at lag %d; rms exp %.3g got %.3g" % (best[0], best[1], np.std(exp[w0:]), np.std(y[w0:])))
