import sys, time, numpy as np, torch
sys.path.insert(0,'.')
import bench, cutesdr_b200 as cs
in_rate, nch, modes, carriers, infos, audio_rate, desc = bench.channel_plan("cfg4", 0, 1)
bank = cs.ReceiverBank(nch, in_rate)
bank.SetAudioRate(audio_rate)
for c in range(nch):
    bank.SetDemod(c, modes[c], infos[c]); bank.SetDemodFreq(c, -carriers[c])
L = bank.block_length()
x = torch.randn(4, 2*L, device='cuda')*1000
aud = torch.zeros(nch, 2304, device='cuda')
for i in range(10): bank.process_device(x[i%4].data_ptr(), L, aud.data_ptr(), 2304)
bank.synchronize()
t0=time.perf_counter()
for i in range(200): bank.process_device(x[i%4].data_ptr(), L, aud.data_ptr(), 2304)
t1=time.perf_counter()
bank.synchronize()
t2=time.perf_counter()
print("enqueue per step %.3f ms, total per step %.3f ms"%((t1-t0)/200*1e3,(t2-t0)/200*1e3))
