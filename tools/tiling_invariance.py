"""Debug aid: is the decimator output bit-identical for different kernel-1 tilings?"""
import os, sys
import numpy as np
sys.path.insert(0, '.')
import cutesdr_b200 as cs
from cutesdr_b200 import modes as M

def run(fs, mode, tile, nblk=3, nohb=False):
    os.environ["CUTESDR_TILE"] = str(tile)
    if nohb: os.environ["CUTESDR_NO_HBCHAIN"] = "1"
    else: os.environ.pop("CUTESDR_NO_HBCHAIN", None)
    bank = cs.ReceiverBank(2, fs)
    for c in range(2):
        bank.SetDemod(c, mode, M.demod_info(mode, HiCut=2800, LowCut=100) if mode == M.DEMOD_USB else M.demod_info(mode))
        bank.SetDemodFreq(c, -1.0e6 * (c + 1))
    bank.tap_enable(1, (1, 2))
    L = bank.block_length()
    rng = np.random.default_rng(3)
    iq = (1000 * (rng.standard_normal(nblk * L) + 1j * rng.standard_normal(nblk * L))).astype(np.complex64)
    bank.ProcessData(iq)
    return bank.tap_read(1, 1), bank.tap_read(1, 2)

for fs, mode in ((20e6, M.DEMOD_USB), (100147200.0, M.DEMOD_FM)):
    a1, a2 = run(fs, mode, 4096)
    for tile, nohb in ((2048, False), (8192, False), (4096, True)):
        b1, b2 = run(fs, mode, tile, nohb=nohb)
        d = np.abs(a1 - b1)
        print(fs, mode, "tile", tile, "nohbchain", nohb, "tap1 equal:", np.array_equal(a1, b1), "max diff", d.max(), "first diff idx", int(np.argmax(d > 0)) if d.max() > 0 else -1, "of", len(a1))
