"""UDP packet front end of the bank (CUdpThread::OnreadyRead, interface/netiobase.cpp:464-534): the 4-byte headers are
stripped and the sequence gaps counted inside cutesdr_bank_process_packets; the payload keeps its wire format down to
kernel 1's tile load."""
import numpy as np
import pytest

import cutesdr_b200 as cs
from cutesdr_b200 import modes as M
from cutesdr_b200.synth import carrier_grid, syn_iq

pytestmark = pytest.mark.gpu


def _missed_packets_reference(seqs):
    """m_MissedPackets as interface/netiobase.cpp:484-496 computes it (qint16 difference, 0 restarts, wrap skips 0)"""
    last, missed = 0, 0
    for s in seqs:
        if s == 0:
            last = 0
        if s != last:
            missed += int(np.int16(np.uint16(s))) - int(np.int16(np.uint16(last)))
            last = s
        last = (last + 1) & 0xFFFF
        if last == 0:
            last = 1
    return missed


@pytest.mark.parametrize("bits", [16, 24])
def test_packet_ingest_strips_headers_and_counts_gaps(bits):
    fs, nch = 2e6, 4
    modes = [[M.DEMOD_AM, M.DEMOD_USB][c % 2] for c in range(nch)]
    carriers = carrier_grid(nch, 200e3)
    infos = [M.demod_info(m, HiCut=2800, LowCut=100) if m == M.DEMOD_USB else M.demod_info(m) for m in modes]

    def make():
        b = cs.ReceiverBank(nch, fs)
        for c in range(nch):
            b.SetDemod(c, modes[c], infos[c])
            b.SetDemodFreq(c, -carriers[c])
        return b

    per = 256 if bits == 16 else 240
    pkt_bytes = 1028 if bits == 16 else 1444
    npk = 700
    n = npk * per
    base = syn_iq(fs, n, modes, carriers, seed=77)
    base = base * np.float32(30000.0 / max(np.abs(base.real).max(), np.abs(base.imag).max()))
    if bits == 16:
        v = np.round(np.stack([base.real, base.imag], axis=1)).astype("<i2")
        payload = v.view(np.uint8).reshape(npk, per * 4)
        raw = v
    else:
        v24 = np.round(np.stack([base.real, base.imag], axis=1) * 256.0).astype(np.int32)
        p = np.empty((n, 2, 3), dtype=np.uint8)
        p[:, :, 0] = v24 & 0xff
        p[:, :, 1] = (v24 >> 8) & 0xff
        p[:, :, 2] = (v24 >> 16) & 0xff
        payload = p.reshape(npk, per * 6)
        raw = p.reshape(-1)
    # sequence numbers: start at 0, run up to the 16-bit wrap (which skips 0), lose 3 packets, go on
    seqs, s = [], 0
    for k in range(npk):
        if k == 0:
            s = 0
        seqs.append(s)
        s = (s + 1) & 0xFFFF
        if s == 0:
            s = 1
        if k == 200:
            s = 65500               # the radio's counter jumps ahead (packets lost)
        if k == 400:
            s = (s + 3) & 0xFFFF    # three packets lost
    packets = np.zeros((npk, pkt_bytes), dtype=np.uint8)
    packets[:, 0] = 0x04
    packets[:, 1] = 0x84 if bits == 16 else 0xA4
    packets[:, 2] = np.array(seqs) & 0xff
    packets[:, 3] = np.array(seqs) >> 8
    packets[:, 4:] = payload
    a_ref, n_ref = make().ProcessRaw(raw, 1 if bits == 16 else 2)
    bank = make()
    # two calls: the sequence state and the partially filled DSP block carry over
    a1, n1 = bank.ProcessPackets(packets[:333].reshape(-1), pkt_bytes, audio_stride=a_ref.shape[1])
    a2, n2 = bank.ProcessPackets(packets[333:].reshape(-1), pkt_bytes, audio_stride=a_ref.shape[1])
    assert np.array_equal(n_ref, n1 + n2) and n_ref.max() >= 3 * 1024
    for c in range(nch):
        got = np.concatenate([a1[c, :n1[c]], a2[c, :n2[c]]])
        assert np.array_equal(a_ref[c, :n_ref[c]], got)
    want = _missed_packets_reference(seqs)
    assert want != 0 and bank.MissedPackets() == want
    assert bank.MissedPackets(reset=True) == want and bank.MissedPackets() == 0
    with pytest.raises(Exception):
        bank.ProcessPackets(np.zeros(1000, dtype=np.uint8), 1000)
