"""Debug aid (GPU box): SAM / FM audio of the GPU bank vs the oracle from the FIRST sample, SNR per 1024-sample burst.
Shows how far the acquisition transient agrees now that the start-up burst is computed in direct form."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cutesdr_b200 as cs
from cutesdr_b200 import modes as M
from cutesdr_b200.synth import snr_db, syn_iq
from oracle import oracle_binding as ob
from oracle import ref_binding as rb

fs, fc, n = 2e6, 250000.0, 700000
for mode in (M.DEMOD_SAM, M.DEMOD_FM, M.DEMOD_AM):
    iq = syn_iq(fs, n, [mode], [fc], seed=20261, total_amp=8000.0)
    info = M.demod_info(mode)
    outs = {}
    for name, mk in (("oracle", ob.Demodulator), ("ref", rb.RefDemodulator)):
        a = mk()
        a.SetInputSampleRate(fs)
        a.SetDemod(mode, info)
        a.SetDemodFreq(-fc)
        outs[name] = a.run(iq)
    bank = cs.ReceiverBank(1, fs)
    bank.SetDemod(0, mode, info)
    bank.SetDemodFreq(0, -fc)
    audio, n_out = bank.ProcessData(iq)
    yb = audio[0, :n_out[0]].astype(np.float64)
    ya, yr = outs["oracle"], outs["ref"]
    print(M.MODE_NAMES[mode], "len", len(ya), len(yr), len(yb), "total SNR gpu-vs-ref %.1f  gpu-vs-oracle %.1f  oracle-vs-ref %.1f" % (
        snr_db(yr, yb), snr_db(ya, yb), snr_db(yr, ya)))
    for k in range(min(10, len(ya) // 1024)):
        s = slice(k * 1024, (k + 1) * 1024)
        print("   burst %d: gpu-vs-ref %.1f dB  oracle-vs-ref %.1f dB   rms ref %.3g" % (k, snr_db(yr[s], yb[s]), snr_db(yr[s], ya[s]),
                                                                                   np.sqrt(np.mean(yr[s] ** 2))))
    d = np.abs(yr - yb)
    print("   first 16 ref:", np.array2string(yr[:8], precision=4), " gpu:", np.array2string(yb[:8], precision=4))
