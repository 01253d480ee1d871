"""Kernel 1T (tensor cores) vs kernel 1 (CUDA cores) vs the C oracle on one down-converter (debug aid)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import cutesdr_b200 as cs
from cutesdr_b200.synth import snr_db
from oracle import oracle_binding as ob

fs = float(sys.argv[1]) if len(sys.argv) > 1 else 100147200.0
bw = 5000.0
nblk = 3
rng = np.random.default_rng(3)
dc0 = cs.CDownConvert(); rate = dc0.SetDataRate(fs, bw); stages = dc0.stages()
dec = 1 << len(stages)
L = int(fs / 100) // dec * dec
L = L // 256 * 256 // dec * dec
print("stages", stages, "rate", rate, "L", L)
n = nblk * L
t = np.arange(n)
x = (3000 * np.exp(2j * np.pi * (1.2345e6 + 700.0) / fs * t) + 8000 * np.exp(2j * np.pi * (-7.1e6) / fs * t)
     + 200 * (rng.standard_normal(n) + 1j * rng.standard_normal(n))).astype(np.complex64)
outs = {}
for name, env in (("tc", None), ("legacy", "1")):
    if env: os.environ["CUTESDR_NO_TC"] = env
    else: os.environ.pop("CUTESDR_NO_TC", None)
    d = cs.CDownConvert(); d.SetDataRate(fs, bw); d.SetFrequency(-1.2345e6)
    t0 = time.time()
    outs[name] = np.concatenate([d.ProcessData(x[k * L:(k + 1) * L]) for k in range(nblk)])
    print(name, "n_out", len(outs[name]), "%.2fs" % (time.time() - t0))
o = ob.DownConvert(); o.SetDataRate(fs, bw); o.SetFrequency(-1.2345e6)
ref = np.concatenate([o.ProcessData(x[k * L:(k + 1) * L]) for k in range(nblk)])
skip = 8
for a in ("tc", "legacy"):
    print("%s vs oracle: %.1f dB" % (a, snr_db(ref[skip:], outs[a][skip:])))
print("tc vs legacy: %.1f dB" % snr_db(outs["legacy"][skip:], outs["tc"][skip:]))
print("first samples tc    ", outs["tc"][:3])
print("first samples legacy", outs["legacy"][:3])
print("first samples oracle", ref[:3])
