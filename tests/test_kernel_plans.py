"""CPU checks of the index algebra two kernels rely on (numpy models of the device code, no GPU):

* k_fastfir (cutesdr_b200/csrc/fastfir.cu): the 2048-point transform as three Stockham passes of radix 16, 16, 8 with the
  output slots of the in-register butterflies, and the fact that the last forward pass leaves thread j with exactly the
  inputs of the first inverse pass;
* k_mix_tc's fp16 form (cutesdr_b200/csrc/decimator.cu): an int16 sample splits exactly into two fp16 numbers by the
  bit patterns the producer warps build.
"""
import numpy as np

N = 2048


def _dft4(a0, a1, a2, a3, conj):
    rot = (lambda z: 1j * z) if conj else (lambda z: -1j * z)
    s02, d02, s13, r13 = a0 + a2, a0 - a2, a1 + a3, rot(a1 - a3)
    return s02 + s13, d02 + r13, s02 - s13, d02 - r13


def _w16(m, conj):
    return np.exp((2j if conj else -2j) * np.pi * m / 16)


def _dft16(a, conj):
    a = list(a)
    for r0 in range(4):
        a[r0], a[r0 + 4], a[r0 + 8], a[r0 + 12] = _dft4(a[r0], a[r0 + 4], a[r0 + 8], a[r0 + 12], conj)
    for i, m in ((5, 1), (6, 2), (7, 3), (9, 2), (10, 4), (11, 6), (13, 3), (14, 6), (15, 9)):
        a[i] = a[i] * _w16(m, conj)
    for q0 in range(4):
        a[4 * q0], a[4 * q0 + 1], a[4 * q0 + 2], a[4 * q0 + 3] = _dft4(a[4 * q0], a[4 * q0 + 1], a[4 * q0 + 2], a[4 * q0 + 3], conj)
    return a


def _dft8(a, conj):
    a = list(a)
    for r0 in range(4):
        a[r0], a[r0 + 4] = a[r0] + a[r0 + 4], a[r0] - a[r0 + 4]
    a[5] *= _w16(2, conj)
    a[6] *= _w16(4, conj)
    a[7] *= _w16(6, conj)
    a[0], a[1], a[2], a[3] = _dft4(a[0], a[1], a[2], a[3], conj)
    a[4], a[5], a[6], a[7] = _dft4(a[4], a[5], a[6], a[7], conj)
    return a


def _slot16(q):
    return 4 * (q & 3) + (q >> 2)


def _slot8(q):
    return 4 * (q & 1) + (q >> 1)


def _tw(m, conj):
    return np.exp((2j if conj else -2j) * np.pi * m / N)


def _fft2048(thread_regs, conj):
    """thread_regs[j][m] = element j + 128 m (what thread j holds on entry); returns the same layout of the transform."""
    buf = np.zeros(N, complex)
    for j in range(128):
        a = _dft16(thread_regs[j], conj)
        for q in range(16):
            buf[16 * j + q] = a[_slot16(q)]
    buf2 = np.zeros(N, complex)
    for j in range(128):
        k = j & 15
        a = _dft16([buf[j + 128 * r] * _tw(8 * k * r, conj) for r in range(16)], conj)
        o = ((j - k) << 4) + k
        for q in range(16):
            buf2[o + 16 * q] = a[_slot16(q)]
    out = [[None] * 16 for _ in range(128)]
    for j in range(128):
        for h in range(2):
            jj = j + 128 * h
            a = _dft8([buf2[jj + 256 * r] * _tw(jj * r, conj) for r in range(8)], conj)
            for q in range(8):
                out[j][h + 2 * q] = a[_slot8(q)]           # element jj + 256 q = j + 128 (h + 2 q)
    return out


def test_fastfir_register_pass_plan_is_a_dft_and_chains_into_its_inverse():
    rng = np.random.default_rng(3)
    x = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    regs = [[x[j + 128 * m] for m in range(16)] for j in range(128)]
    X = _fft2048(regs, False)
    ref = np.fft.fft(x)
    got = np.array([X[j][m] for m in range(16) for j in range(128)])       # element j + 128 m
    assert np.abs(got - ref).max() < 1e-9
    # the forward transform's output registers ARE the inverse transform's input registers (no exchange in between)
    y = _fft2048(X, True)
    back = np.array([y[j][m] for m in range(16) for j in range(128)]) / N
    assert np.abs(back - x).max() < 1e-12


def test_fastfir_shared_memory_padding_is_conflict_free():
    """element i lives at i + i / 16; every access pattern of the passes maps a half warp onto 16 distinct 8-byte banks"""
    def banks(idx):
        return [((i + (i >> 4)) * 2) % 32 for i in idx]
    for w in range(4):
        lanes = range(32 * w, 32 * w + 32)
        for q in (0, 7, 15):
            for half in (0, 16):
                ls = list(lanes)[half:half + 16]
                assert len(set(banks([16 * j + q for j in ls]))) == 16                 # pass-1 stores
                assert len(set(banks([j + 128 * q for j in ls]))) == 16                # pass-2 loads
                assert len(set(banks([((j - (j & 15)) << 4) + (j & 15) + 16 * q for j in ls]))) == 16    # pass-2 stores
                assert len(set(banks([j + 256 * (q & 7) for j in ls]))) == 16          # pass-3 loads


def test_int16_splits_exactly_into_two_fp16_numbers():
    v = np.arange(-32768, 32768, dtype=np.int64)
    u = (v + 32768).astype(np.uint32)
    hi_field, lo_field = (u >> 6) & 0x3FF, u & 0x3F
    # the bit patterns the producer warps build: 0x6400 | field is the fp16 number 1024 + field
    hi_h = (0x6400 | hi_field).astype(np.uint16).view(np.float16)
    lo_h = (0x6400 | lo_field).astype(np.uint16).view(np.float16)
    assert np.array_equal(hi_h.astype(np.float64), 1024.0 + hi_field)
    assert np.array_equal(lo_h.astype(np.float64), 1024.0 + lo_field)
    hi = (hi_h - np.float16(1536.0)).astype(np.float16)                                  # HADD2
    lo = (lo_h.astype(np.float64) * (1.0 / 64.0) - 16.0).astype(np.float16)             # HFMA2: one rounding, exact here
    assert np.array_equal(hi.astype(np.float64), np.floor((v + 32768) / 64.0) - 512.0)
    assert np.array_equal(lo.astype(np.float64), ((v + 32768) % 64) / 64.0)
    assert np.array_equal(hi.astype(np.float64) + lo.astype(np.float64), v / 64.0)
    # the float path used for halo samples and boundary tiles gives the same two numbers
    x = v.astype(np.float32)
    uu = x + np.float32(32768.0)
    q = np.floor(uu * np.float32(1.0 / 64.0))
    assert np.array_equal((q - np.float32(512.0)).astype(np.float16), hi)
    assert np.array_equal(((uu - np.float32(64.0) * q) * np.float32(1.0 / 64.0)).astype(np.float16), lo)


def test_sequential_kernel_tile_swizzle_is_conflict_free():
    """post.cu tiles: 32 rows x 16 doubles, 16-byte chunk j of row r at chunk j ^ (r & 7). A quarter warp's 16-byte accesses
    must cover all 32 banks both for the per-lane row accesses and for the cooperative 8-lanes-per-row copies."""
    def bank16(row, chunk):            # 16-byte bank group (8 of them = 128 bytes)
        return ((row * 8) + (chunk ^ (row & 7))) % 8
    for j in range(8):
        for quarter in range(4):
            rows = range(8 * quarter, 8 * quarter + 8)
            assert len({bank16(r, j) for r in rows}) == 8                  # lane = row, same logical chunk j
    for i in range(8):
        for quarter in range(4):
            lanes = range(8 * quarter, 8 * quarter + 8)
            acc = {bank16(4 * i + (lane >> 3), lane & 7) for lane in lanes}        # 8 lanes cover one row segment
            assert len(acc) == 8
