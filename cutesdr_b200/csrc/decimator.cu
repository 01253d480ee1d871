// decimator.cu -- kernel 1 (fused NCO mix + CIC3 cascade) and kernel 2 (half-band stages).
//
// Data layout
//   wideband block x : complex64[L] in HBM, kHaloMax samples of the previous block in front.
//   K1  : grid (time tiles, channel blocks). A CTA stages one time tile of x in shared memory
//         (float4 loads); every LANE is one CHANNEL and walks the tile sequentially, so the
//         tile is read with broadcast LDS, the oscillator, the mixer product and all CIC3
//         states live in registers, and nothing is exchanged between lanes. Tiles overlap by
//         a halo that re-primes the (feed-forward) CIC states, so tiles are independent.
//         Output (rate fs/2^ncic) goes time-major/channel-minor [row][stride]: one coalesced
//         256-byte store per warp and output row.
//   K2  : one launch per half-band stage, thread per (output row, channel), taps in constant
//         memory, symmetric taps folded; each stage's input is a power-of-two ring of rows so
//         the N-1 rows of history simply persist between blocks. The last stage transposes
//         into the per-channel ring [c][kDecRing] that CFastFIR's overlap-save windows read.
//
// Reference: CDownConvert::ProcessData and the three DecBy2 classes, dsp/downconvert.cpp:186-460.
#include "decimator.cuh"
#include "halfband_tables.h"
#include <cuda.h>            // CUtensorMap + cuTensorMapEncodeTiled prototype (resolved at run time through the runtime)
#include <cuda_fp16.h>
#include <stdlib.h>

namespace csdr {

// ------------------------------------------------------------------------------------------
// host: stage ladder (dsp/downconvert.cpp:127-166)
// ------------------------------------------------------------------------------------------
Tuning Tuning::from_env()
{
    Tuning t;
    t.no_tc = getenv("CUTESDR_NO_TC") != nullptr;
    t.no_hbchain = getenv("CUTESDR_NO_HBCHAIN") != nullptr;
    t.hbtail = getenv("CUTESDR_HBTAIL") != nullptr;
    t.no_overlap = getenv("CUTESDR_NO_OVERLAP") != nullptr;
    t.debug_timing = getenv("CUTESDR_DEBUG_TIMING") != nullptr;
    if (const char* e = getenv("CUTESDR_FUSE_HB")) t.fuse_hb = atoi(e);
    if (const char* e = getenv("CUTESDR_TILE")) t.tile = atoi(e);
    if (const char* e = getenv("CUTESDR_TC_SEG")) t.tc_seg = atoi(e);
    t.no_hbstream = getenv("CUTESDR_NO_HBSTREAM") != nullptr;
    t.no_hbtail = getenv("CUTESDR_NO_HBTAIL") != nullptr;
    if (const char* e = getenv("CUTESDR_HS_HALO")) t.hs_halo = atoi(e);
    if (const char* e = getenv("CUTESDR_TC_SPARE")) t.tc_spare = std::max(0, atoi(e));
    if (const char* e = getenv("CUTESDR_HS_CTAS")) t.hs_ctas = std::max(1, atoi(e));
    return t;
}

double plan_stages(double in_rate, double max_bw, std::vector<int>& lens)
{
    lens.clear();
    double f = in_rate;
    const double last_limit = .5 - csdr_hb_alias_free[CSDR_HB_NUM_KINDS - 1];
    while (f > (max_bw / last_limit) && f > (7900.0 * 2.0)) {      // MIN_OUTPUT_RATE, :52
        for (int k = 0; k < CSDR_HB_NUM_KINDS; k++) {
            if (f >= (max_bw / (.5 - csdr_hb_alias_free[k]))) {
                lens.push_back(csdr_hb_len[k]);
                break;
            }
        }
        f /= 2.0;
    }
    return f;
}

// ------------------------------------------------------------------------------------------
// device helpers
// ------------------------------------------------------------------------------------------
__constant__ float c_hb_taps[88];
__constant__ float c_cic4[48];      // impulse response of four cascaded CIC3 decimate-by-2 stages (46 taps, unnormalised)
static std::once_flag g_taps_once[16];

static int upload_taps()
{
    int dev = 0;
    CSDR_CK(cudaGetDevice(&dev));
    cudaError_t err = cudaSuccess;
    std::call_once(g_taps_once[dev & 15], [&]() {
        float h[88];
        for (int i = 0; i < 88; i++) h[i] = (float)csdr_hb_taps[i];
        err = cudaMemcpyToSymbol(c_hb_taps, h, sizeof(h));
        if (err != cudaSuccess) return;
        // four cascaded CIC3 decimate-by-2 stages as one FIR: (1,3,3,1) upsampled by 1, 2, 4, 8 and convolved
        std::vector<double> acc(1, 1.0);
        for (int up = 1; up <= 8; up <<= 1) {
            std::vector<double> nxt(acc.size() + 3 * up, 0.0);
            const double k[4] = {1.0, 3.0, 3.0, 1.0};
            for (size_t a = 0; a < acc.size(); a++)
                for (int b = 0; b < 4; b++) nxt[a + (size_t)b * up] += acc[a] * k[b];
            acc.swap(nxt);
        }
        float c4[48] = {0};
        for (size_t a = 0; a < acc.size() && a < 48; a++) c4[a] = (float)acc[a];       // 46 taps, integers < 2^24
        err = cudaMemcpyToSymbol(c_cic4, c4, sizeof(c4));
    });
    if (err != cudaSuccess) { set_error("tap upload: %s", cudaGetErrorString(err)); return CUTESDR_E_CUDA; }
    return CUTESDR_OK;
}

struct OutDesc {
    float2* p;
    unsigned mask;        // rows-1 (time-major ring)
    int stride;           // channels per row
    int transposed;       // 1: per-channel ring [c][kDecRing]
    long long base;       // absolute row index of this block's first output
};

__device__ __forceinline__ void store_out(const OutDesc& od, long long row, int c, float2 v)
{
    if (od.transposed)
        od.p[(size_t)c * kDecRing + (size_t)((od.base + row) & (kDecRing - 1))] = v;
    else
        od.p[(size_t)((od.base + row) & od.mask) * od.stride + c] = v;
}

// Explicit rounding/contraction: every inlined copy must round identically, otherwise the oscillator
// (and with it every output bit) would depend on which code path produced it.
__device__ __forceinline__ float2 cmul(float2 a, float2 b)
{
    return make_float2(__fmaf_rn(a.x, b.x, -__fmul_rn(a.y, b.y)), __fmaf_rn(a.x, b.y, __fmul_rn(a.y, b.x)));
}

// exact oscillator seed from the 64-bit phase (turns * 2^64): angle = pi * (hi32 / 2^31)
__device__ __forceinline__ float2 seed_osc(unsigned long long ph)
{
    float t = (float)(int)(unsigned)(ph >> 32) * (1.0f / 2147483648.0f);
    float s, c;
    sincospif(t, &s, &c);
    return make_float2(c, s);
}

constexpr int kFuseHb = 1;      // a second fused stage costs kernel 1 more (registers, halo) than it saves in kernel 2
constexpr int kFuseHbTc = 2;    // kernel 1T: every fused stage halves its HBM output and takes a pass off kernel 2

static int k1_body(int ncic) { int g = 1 << ncic; return g < 32 ? 32 : g; }
static int k1_halo(int ncic, int nhb)
{
    const int g = 1 << ncic, b = k1_body(ncic);
    const int need = (ncic == 0 ? 0 : (2 << ncic)) + 10 * g * ((1 << nhb) - 1);
    const int q = std::max(g << nhb, b);
    return (need + q - 1) / q * q;
}

// Wideband input formats kernel 1 reads directly (the radio's wire formats, interface/netiobase.cpp:497-527):
//   0 = complex64;  1 = interleaved little-endian int16 I,Q (value = the integer);
//   2 = packed little-endian int24 I,Q (value = integer / 256: +-32768 range with 8 fraction bits).
// The conversion to float32 is exact for all three.
__device__ __forceinline__ float2 fetch_sample(const void* __restrict__ x, int fmt, int j)
{
    if (fmt == 0) return reinterpret_cast<const float2*>(x)[j];
    if (fmt == 1) {
        const short2 v = reinterpret_cast<const short2*>(x)[j];
        return make_float2((float)v.x, (float)v.y);
    }
    const unsigned char* p = reinterpret_cast<const unsigned char*>(x) + (size_t)6 * j;
    const int i24 = (int)((unsigned)p[0] << 8 | (unsigned)p[1] << 16 | (unsigned)p[2] << 24) >> 8;
    const int q24 = (int)((unsigned)p[3] << 8 | (unsigned)p[4] << 16 | (unsigned)p[5] << 24) >> 8;
    return make_float2((float)i24 * (1.0f / 256.0f), (float)q24 * (1.0f / 256.0f));
}

// The last kHaloMax samples of [previous halo | this block] become the next block's halo
// (written to the other half of a double buffer, so concurrent readers of halo_cur are safe).
__device__ __forceinline__ void save_halo(const void* __restrict__ x, int fmt, const float2* __restrict__ halo_cur,
                                          float2* __restrict__ halo_next, int L)
{
    for (int i = threadIdx.x; i < kHaloMax; i += blockDim.x) {
        const int j = L - kHaloMax + i;
        halo_next[i] = j >= 0 ? fetch_sample(x, fmt, j) : halo_cur[kHaloMax + j];
    }
}

struct CicSt { float2 xodd, xeven; };
struct Hb11St { float2 e[5]; float2 o[3]; };     // e[0] = most recent even-indexed input, o[0] = most recent odd

// NCIC CIC3 stages followed by NHB fused 11-tap half-bands
template <int NCIC, int NHB> struct K1Cfg {
    static constexpr int G = 1 << NCIC;                    // input samples per CIC-chain output
    static constexpr int B = G < 32 ? 32 : G;              // unrolled body length
    static constexpr int Hcic = NCIC == 0 ? 0 : (2 << NCIC);                  // >= 2^(ncic+1)-2 samples re-prime the CICs
    static constexpr int Hneed = Hcic + 10 * G * ((1 << NHB) - 1);            // + 10 inputs of history per half-band
    static constexpr int Q = (G << NHB) > B ? (G << NHB) : B;                 // tiles start on output boundaries
    static constexpr int H = (Hneed + Q - 1) / Q * Q;
};

struct EmitCtx {
    OutDesc od;
    long long row_lo;     // first output row this tile owns (rows before it belong to the halo)
    long long row_hi;     // one past the last row this tile owns (kernel 1T's last MMA tile overshoots)
    int c;
    float scale;
    float h0, h2, h4;     // HB11 taps 0/2/4 (= 10/8/6); centre tap is 0.5
    __device__ __forceinline__ void emit(long long q, float2 v) const
    {
        if (q >= row_lo && q < row_hi) store_out(od, q, c, make_float2(v.x * scale, v.y * scale));
    }
};

// Kernel 1T's emitter: its outputs leave strictly in row order, so the ring position and the
// owned-range test are running 32-bit counters instead of 64-bit index arithmetic per store.
struct TcEmit {
    float2* p0;           // ring base + this channel's offset
    unsigned mask;        // ring rows - 1
    unsigned pos;         // ring row of the next output
    int estride;          // float2 elements between ring rows
    int rel;              // next output's row - first owned row (negative while priming)
    int n_rows;           // rows this segment owns
    float scale;
    float h0, h2, h4;
    __device__ __forceinline__ void emit(long long, float2 v)
    {
        float2* dst = p0 + (size_t)pos * (size_t)estride;
        const unsigned ok = (unsigned)rel < (unsigned)n_rows;
        // predicated store, no branch: priming outputs and the overshoot of the last tile are dropped
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.global.v2.f32 [%1], {%2, %3};\n\t}\n" ::"r"(ok), "l"(dst),
                     "f"(v.x * scale), "f"(v.y * scale)
                     : "memory");
        rel++;
        pos = (pos + 1) & mask;
    }
};

// 11-tap half-band decimate-by-2 in direct form on a register delay line
// (y[m] = sum_j h[j] x[2m-10+j], dsp/downconvert.cpp:348-423): an output is complete when the
// even-indexed input x[2m] arrives; odd-indexed inputs only feed the centre tap three outputs later.
// q = block-relative index of v at this stage (may be negative inside the halo).
template <int NHB, int J, class E>
__device__ __forceinline__ void hb_feed(float2 v, long long q, Hb11St* hs, E& em)
{
    if constexpr (J == NHB) {
        em.emit(q, v);
    } else {
        Hb11St& s = hs[J];
        if ((q & 1) == 0) {
            float2 y;
            y.x = fmaf(em.h0, s.e[4].x + v.x, fmaf(em.h2, s.e[3].x + s.e[0].x, fmaf(em.h4, s.e[2].x + s.e[1].x, 0.5f * s.o[2].x)));
            y.y = fmaf(em.h0, s.e[4].y + v.y, fmaf(em.h2, s.e[3].y + s.e[0].y, fmaf(em.h4, s.e[2].y + s.e[1].y, 0.5f * s.o[2].y)));
            s.e[4] = s.e[3]; s.e[3] = s.e[2]; s.e[2] = s.e[1]; s.e[1] = s.e[0]; s.e[0] = v;
            hb_feed<NHB, J + 1, E>(y, q >> 1, hs, em);
        } else {
            s.o[2] = s.o[1]; s.o[1] = s.o[0]; s.o[0] = v;
        }
    }
}

// First fused half-band when a body carries an EVEN number of CIC outputs (B/G >= 2): tiles start on
// multiples of 2G, so the body's CIC output IDX is even-indexed iff IDX is even -- a compile-time
// fact, no branch, and the compiler can keep the delay line in fixed registers.
template <int NHB, int IDX, class E>
__device__ __forceinline__ void hb_feed_static(float2 v, int q0, Hb11St* hs, E& em)
{
    Hb11St& s = hs[0];
    if constexpr ((IDX & 1) == 0) {
        float2 y;
        y.x = fmaf(em.h0, s.e[4].x + v.x, fmaf(em.h2, s.e[3].x + s.e[0].x, fmaf(em.h4, s.e[2].x + s.e[1].x, 0.5f * s.o[2].x)));
        y.y = fmaf(em.h0, s.e[4].y + v.y, fmaf(em.h2, s.e[3].y + s.e[0].y, fmaf(em.h4, s.e[2].y + s.e[1].y, 0.5f * s.o[2].y)));
        s.e[4] = s.e[3]; s.e[3] = s.e[2]; s.e[2] = s.e[1]; s.e[1] = s.e[0]; s.e[0] = v;
        hb_feed<NHB, 1, E>(y, (long long)((q0 + IDX) >> 1), hs, em);
    } else {
        s.o[2] = s.o[1]; s.o[1] = s.o[0]; s.o[0] = v;
    }
}

// CIC3 decimate-by-2, scale .125 folded into the kernel's output scale
// (y = odd + Xeven + 3*(Xodd + even), dsp/downconvert.cpp:450-455).
template <int NCIC, int NHB, int S, int IDX, class E>
__device__ __forceinline__ void cic_feed(float2 v, CicSt* st, float2* ev, Hb11St* hs, long long q0, E& em)
{
    if constexpr (S == NCIC) {
        constexpr int G = 1 << NCIC, B = G < 32 ? 32 : G;
        if constexpr (NHB >= 1 && (B / G) >= 2) hb_feed_static<NHB, IDX, E>(v, (int)q0, hs, em);
        else hb_feed<NHB, 0, E>(v, q0 + IDX, hs, em);
    } else if constexpr ((IDX & 1) == 0) {
        ev[S] = v;
    } else {
        float2 e = ev[S], r;
        r.x = (v.x + st[S].xeven.x) + 3.0f * (st[S].xodd.x + e.x);
        r.y = (v.y + st[S].xeven.y) + 3.0f * (st[S].xodd.y + e.y);
        st[S].xodd = v;
        st[S].xeven = e;
        cic_feed<NCIC, NHB, S + 1, (IDX >> 1), E>(r, st, ev, hs, q0, em);
    }
}

template <int NCIC, int NHB, int K, int B> struct Body {
    static __device__ __forceinline__ void run(const float2* t, float2& o, float2 w, CicSt* st, float2* ev, Hb11St* hs,
                                               long long q0, const EmitCtx& em)
    {
        float2 xv = t[K];                 // all lanes read the same address: broadcast LDS
        float2 y = cmul(xv, o);           // mixer, dsp/downconvert.cpp:238-239
        if (K + 1 < B) o = cmul(o, w);    // oscillator step, :211-212
        cic_feed<NCIC, NHB, 0, K, const EmitCtx>(y, st, ev, hs, q0, em);
        Body<NCIC, NHB, K + 1, B>::run(t, o, w, st, ev, hs, q0, em);
    }
};
template <int NCIC, int NHB, int B> struct Body<NCIC, NHB, B, B> {
    static __device__ __forceinline__ void run(const float2*, float2&, float2, CicSt*, float2*, Hb11St*, long long,
                                               const EmitCtx&) {}
};

// ------------------------------------------------------------------------------------------
// K1: fused NCO mix + NCIC x CIC3 + NHB x HB11
// ------------------------------------------------------------------------------------------
template <int NCIC, int NHB>
__global__ void __launch_bounds__(256, 3) k_mix_cic(const void* __restrict__ x, int fmt, const float2* __restrict__ halo_cur,
                                                 float2* __restrict__ halo_next, int L, int tile_len,
                                                 const NcoDev* __restrict__ nco,
                                                 const unsigned long long* __restrict__ phase_cur,
                                                 unsigned long long* __restrict__ phase_next, int nch,
                                                 OutDesc od, float scale)
{
    typedef K1Cfg<NCIC, NHB> Cfg;
    constexpr int G = Cfg::G, B = Cfg::B, H = Cfg::H;
    extern __shared__ float4 smem4[];
    const int t0 = blockIdx.x * tile_len;
    const int n_tile = min(tile_len, L - t0);
    const int n_load = n_tile + H;
    {
        // samples before the block start (first tile's halo) come from the saved tail of the
        // previous block: halo_cur[kHaloMax + j] for j < 0. Two samples per iteration; the radio's
        // integer formats are unpacked here, on the way into shared memory (no separate pass).
        const float4* hsrc = reinterpret_cast<const float4*>(halo_cur + (kHaloMax + t0 - H));
        const int n_neg = t0 < H ? ((H - t0) >> 1) : 0;
        if (fmt == 0) {
            const float4* src = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(x) + (t0 - H));
            for (int i = threadIdx.x; i < (n_load >> 1); i += blockDim.x) smem4[i] = i < n_neg ? __ldg(hsrc + i) : __ldg(src + i);
        } else if (fmt == 1) {
            const int2* src = reinterpret_cast<const int2*>(reinterpret_cast<const short2*>(x) + (t0 - H));
            for (int i = threadIdx.x; i < (n_load >> 1); i += blockDim.x) {
                if (i < n_neg) smem4[i] = __ldg(hsrc + i);
                else {
                    const int2 v = __ldg(src + i);
                    smem4[i] = make_float4((float)(short)(v.x & 0xffff), (float)(short)(v.x >> 16), (float)(short)(v.y & 0xffff),
                                           (float)(short)(v.y >> 16));
                }
            }
        } else {
            for (int i = threadIdx.x; i < (n_load >> 1); i += blockDim.x) {
                if (i < n_neg) smem4[i] = __ldg(hsrc + i);
                else {
                    const float2 a = fetch_sample(x, 2, t0 - H + 2 * i), b2 = fetch_sample(x, 2, t0 - H + 2 * i + 1);
                    smem4[i] = make_float4(a.x, a.y, b2.x, b2.y);
                }
            }
        }
    }
    if (blockIdx.x == gridDim.x - 1 && blockIdx.y == 0) save_halo(x, fmt, halo_cur, halo_next, L);
    __syncthreads();
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= nch) return;

    const NcoDev p = nco[c];
    const unsigned long long ph0 = phase_cur[c];
    if (blockIdx.x == 0) phase_next[c] = ph0 + (unsigned long long)L * p.inc;
    // sample i of the block is multiplied by e^{j(P + (i+1) inc)} (the reference rotates first)
    unsigned long long ph = ph0 + (unsigned long long)(long long)(t0 - H + 1) * p.inc;
    const unsigned long long ph_step = p.inc * (unsigned long long)B;
    const float2 w1 = make_float2(p.w1c, p.w1s), wg = make_float2(p.wgc, p.wgs);

    CicSt st[NCIC > 0 ? NCIC : 1];
    float2 ev[NCIC > 0 ? NCIC : 1];
    Hb11St hs[NHB > 0 ? NHB : 1];
#pragma unroll
    for (int s = 0; s < (NCIC > 0 ? NCIC : 1); s++) {
        st[s].xodd = make_float2(0.f, 0.f);
        st[s].xeven = make_float2(0.f, 0.f);
        ev[s] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int s = 0; s < (NHB > 0 ? NHB : 1); s++) {
#pragma unroll
        for (int k = 0; k < 5; k++) hs[s].e[k] = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 3; k++) hs[s].o[k] = make_float2(0.f, 0.f);
    }
    EmitCtx em;
    em.od = od;
    em.row_lo = (long long)(t0 / (G << NHB));
    em.row_hi = 0x7fffffffffffffffLL;
    em.c = c;
    em.scale = scale;
    em.h0 = c_hb_taps[0]; em.h2 = c_hb_taps[1]; em.h4 = c_hb_taps[2];

    const float2* tile = reinterpret_cast<const float2*>(smem4);
    const int nbody = n_load / B;
    // The oscillator is re-seeded exactly (sincospi of the 64-bit phase) at ABSOLUTE body indices that
    // are multiples of 8 (every 256 samples of the block) and rotated in float32 in between, so its
    // values -- and therefore every output bit -- do not depend on how the block is cut into tiles
    // (nor on the number of channels or GPUs that decided the tiling).
    const int body0 = (t0 - H) / B;                                  // exact: B divides t0 and H; may be negative
    const int lead = ((body0 % 8) + 8) % 8;                          // bodies since the last absolute seed point
    float2 S = seed_osc(ph - (unsigned long long)lead * ph_step);
    for (int i = 0; i < lead; i++) S = cmul(S, wg);
    long long q0 = (long long)((t0 - H) / G);          // index of the body's first CIC-chain output (exact: G | t0-H)
    for (int b = 0; b < nbody; b++) {
        if (((body0 + b) & 7) == 0) S = seed_osc(ph);
        float2 o = S;
        Body<NCIC, NHB, 0, B>::run(tile + b * B, o, w1, st, ev, hs, q0, em);
        S = cmul(S, wg);
        ph += ph_step;
        q0 += B / G;
    }
}

// Slow generic path for block lengths that are not a multiple of the unrolled body (single-object
// API with odd sizes). Same math, run-time stage counts, one tile.
__global__ void k_mix_cic_generic(const void* __restrict__ x, int fmt, const float2* __restrict__ halo_cur,
                                  float2* __restrict__ halo_next, int L, int ncic, int nhb, const NcoDev* __restrict__ nco,
                                  const unsigned long long* __restrict__ phase_cur,
                                  unsigned long long* __restrict__ phase_next, int nch, OutDesc od, float scale)
{
    if (blockIdx.x == 0) save_halo(x, fmt, halo_cur, halo_next, L);
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nch) return;
    const int G = 1 << ncic;
    const int q_align = G << nhb;
    int H = (ncic == 0 ? 0 : (2 << ncic)) + 10 * G * ((1 << nhb) - 1);
    H = (H + q_align - 1) / q_align * q_align;
    const NcoDev p = nco[c];
    const unsigned long long ph0 = phase_cur[c];
    phase_next[c] = ph0 + (unsigned long long)L * p.inc;
    const float2 w1 = make_float2(p.w1c, p.w1s);
    const float h0 = c_hb_taps[0], h2 = c_hb_taps[1], h4 = c_hb_taps[2];
    float2 xo[8], xe[8], ev[8];
    int par[8];
    Hb11St hs[2];
    for (int s = 0; s < 8; s++) { xo[s] = xe[s] = ev[s] = make_float2(0.f, 0.f); par[s] = 0; }
    for (int s = 0; s < 2; s++) {
        for (int k = 0; k < 5; k++) hs[s].e[k] = make_float2(0.f, 0.f);
        for (int k = 0; k < 3; k++) hs[s].o[k] = make_float2(0.f, 0.f);
    }
    float2 o = make_float2(1.f, 0.f);
    long long q = -(long long)(H >> ncic);
    for (int i = -H; i < L; i++) {
        if ((i & 31) == 0 || i == -H) o = seed_osc(ph0 + (unsigned long long)(long long)(i + 1) * p.inc);
        float2 v = cmul(i >= 0 ? fetch_sample(x, fmt, i) : halo_cur[kHaloMax + i], o);
        o = cmul(o, w1);
        int s = 0;
        while (s < ncic) {
            if (par[s] == 0) { ev[s] = v; par[s] = 1; break; }
            float2 r;
            r.x = (v.x + xe[s].x) + 3.0f * (xo[s].x + ev[s].x);
            r.y = (v.y + xe[s].y) + 3.0f * (xo[s].y + ev[s].y);
            xo[s] = v; xe[s] = ev[s]; par[s] = 0;
            v = r;
            s++;
        }
        if (s < ncic) continue;
        long long qq = q++;
        int j = 0;
        for (; j < nhb; j++) {
            Hb11St& hh = hs[j];
            if (qq & 1) { hh.o[2] = hh.o[1]; hh.o[1] = hh.o[0]; hh.o[0] = v; break; }
            float2 y;
            y.x = fmaf(h0, hh.e[4].x + v.x, fmaf(h2, hh.e[3].x + hh.e[0].x, fmaf(h4, hh.e[2].x + hh.e[1].x, 0.5f * hh.o[2].x)));
            y.y = fmaf(h0, hh.e[4].y + v.y, fmaf(h2, hh.e[3].y + hh.e[0].y, fmaf(h4, hh.e[2].y + hh.e[1].y, 0.5f * hh.o[2].y)));
            hh.e[4] = hh.e[3]; hh.e[3] = hh.e[2]; hh.e[2] = hh.e[1]; hh.e[1] = hh.e[0]; hh.e[0] = v;
            v = y;
            qq >>= 1;
        }
        if (j == nhb && qq >= 0) store_out(od, qq, c, make_float2(v.x * scale, v.y * scale));
    }
}

typedef void (*K1Fn)(const void*, int, const float2*, float2*, int, int, const NcoDev*, const unsigned long long*,
                     unsigned long long*, int, OutDesc, float);

static K1Fn k1_kernel(int ncic, int nhb)
{
    static const K1Fn table[7][3] = {
        {k_mix_cic<0, 0>, k_mix_cic<0, 1>, k_mix_cic<0, 2>}, {k_mix_cic<1, 0>, k_mix_cic<1, 1>, k_mix_cic<1, 2>},
        {k_mix_cic<2, 0>, k_mix_cic<2, 1>, k_mix_cic<2, 2>}, {k_mix_cic<3, 0>, k_mix_cic<3, 1>, k_mix_cic<3, 2>},
        {k_mix_cic<4, 0>, k_mix_cic<4, 1>, k_mix_cic<4, 2>}, {k_mix_cic<5, 0>, k_mix_cic<5, 1>, k_mix_cic<5, 2>},
        {k_mix_cic<6, 0>, k_mix_cic<6, 1>, k_mix_cic<6, 2>}};
    return table[ncic][nhb];
}


// ------------------------------------------------------------------------------------------
// K1T: kernel 1 on the tensor cores (tcgen05, kind::tf32, operands and accumulators in TMEM).
//
// The NCO mix followed by the first FOUR CIC3 stages is one complex FIR-decimate-by-16 per channel,
//     y4[m] = sum_{j<46} H[j] x[16m+15-j] e^{j phi(16m+15-j)}
//           = e^{j phi(16m+15)} * sum_k A_c[k] x[16(m-2)+k],   A_c[k] = H[47-k] e^{-j 2 pi (47-k) f_c/fs},
// i.e. a GEMM  Y[channel, time] = A[channel, 48] * X[48, time]  against a Hankel matrix of the wideband
// block that every channel shares. A CTA owns 128 channels (the MMA's M); per tile of 32 outputs (512
// input samples) it issues  D_re = Ar Xr - Ai Xi,  D_im = Ar Xi + Ai Xr  as M128 N32 K8 MMAs with every
// operand split into tf32 hi + lo parts (3 products per term: lo*hi, hi*lo, hi*hi; the dropped lo*lo is
// 2^-22 relative), fp32 accumulation in TMEM.
//   * A (per-channel coefficients, planes Ar/Ai x hi/lo, 48 columns each) lives in TMEM for the whole
//     CTA: the MMAs read only B from shared memory (an SS-form MMA of this shape saturates the 128 B/clk
//     shared-memory port with its A re-reads).
//   * X is staged by 4 producer warps (a warp pair per segment, each warp half of a tile's chunks, the global
//     loads of the next tile in flight while the current one is converted) as 4 planes (re/im x hi/lo) in
//     NATURAL time order with zero duplication: 16-byte chunk q (4 samples) of X-row m (16 samples) sits at
//     byte 16 m + P q. K-major, no swizzle, LBO = P, SBO = 128: MMA row n reads row m0-2+n+shift, and the three
//     16-sample shifts of the Hankel matrix are just the descriptor start address + 16 * shift.
//   * 2 independent time segments per CTA, interleaved tile by tile; segment e owns two TMEM accumulator slots
//     (tile parity), an epilogue warp set (4 warps, TMEM lane = channel) and two MMA-issuing warps (one per
//     re / im accumulator, so every accumulator has exactly one writer and a fixed MMA order). An epilogue thread
//     pulls 16 consecutive outputs per tcgen05.ld -- the next 16 in flight under the arithmetic of the current
//     16 --, multiplies by the oscillator at the decimated rate (re-seeded exactly from the 64-bit phase every
//     32 outputs), and runs the remaining CIC3 / fused 11-tap half-band stages and the store exactly as kernel 1
//     does. The accumulator slot is released as soon as the tile is in registers.
//   * mbarriers + tcgen05.commit order producers -> MMA -> epilogue; B runs through an 8-stage ring.
// Segments re-prime the feed-forward stages with a halo of PRE outputs, like kernel 1's tiles.
// ------------------------------------------------------------------------------------------
constexpr int kTcN = 32;                                  // outputs (fs/16) per MMA tile
constexpr int kTcRows = kTcN + 2;                         // X rows per tile (two rows of history)
constexpr int kTcP = kTcRows * 16;                        // bytes per chunk-column q of a plane
constexpr int kTcPlane = 4 * kTcP;                        // bytes per B plane
constexpr int kTcStage = 4 * kTcPlane;                    // Xr_hi, Xr_lo, Xi_hi, Xi_lo
constexpr int kTcStages = 8;
constexpr int kTcBarOff = kTcStages * kTcStage;
constexpr int kTcSmem = kTcBarOff + 512;
constexpr int kTcSegs = 2;                                // time segments (= accumulator slots = epilogue sets) per CTA
constexpr int kTcThreads = 32 * (4 * kTcSegs + 4 + 2 * kTcSegs);    // 8 epilogue + 4 producer + 4 MMA warps
constexpr int kTcAccCol = 192;                            // TMEM: A planes in columns 0..191; segment e, slot b (tile parity) at 192 + 4 kTcN e + 2 kTcN b (re | im)

template <int NCR, int NHB> struct TcCfg {
    static constexpr int Gr = 1 << NCR;                                             // fs/16 samples per CIC-chain output
    static constexpr int need = (NCR == 0 ? 0 : (2 << NCR)) + 10 * Gr * ((1 << NHB) - 1);
    static constexpr int q = (Gr << NHB) > 16 ? (Gr << NHB) : 16;
    static constexpr int PRE = (need + q - 1) / q * q;                              // priming outputs (fs/16) per segment
};
static int tc_pre(int ncr, int nhb)
{
    const int gr = 1 << ncr;
    const int need = (ncr == 0 ? 0 : (2 << ncr)) + 10 * gr * ((1 << nhb) - 1);
    const int q = std::max(gr << nhb, 16);
    return (need + q - 1) / q * q;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float tf32_rn(float x)
{
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_ts_f16(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bar_wait(uint64_t* bar, uint32_t parity)
{
    // try_wait with a suspend-time hint: the waiting warp sleeps in hardware until the phase flips instead of
    // re-issuing the poll (polls share the MIO queue the tcgen05 instructions go through)
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "W_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra D_%=;\n\t"
        "bra W_%=;\n"
        "D_%=:\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(0x989680)
        : "memory");
}
__device__ __forceinline__ void bar_arrive(uint64_t* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp; ptxas recognises elect.sync and keeps the guarded code on the uniform datapath
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
// Compiler-level ordering for values produced by tcgen05.ld: arithmetic on them must stay behind the
// tcgen05.wait::ld that completes the load. Volatile asm statements keep their order, and routing the registers
// through an (empty) one after the wait makes every later use depend on it.
__device__ __forceinline__ void tmem_pin16(float* v)
{
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]), "+f"(v[9]),
                      "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])::"memory");
}
// 16 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v)
{
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float* v)
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
                 "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
                 "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
                 : "memory");
}

// A-operand table: [channel (padded to whole 128-channel groups)][Ar_hi, Ar_lo, Ai_hi, Ai_lo][48]
__global__ void k_tc_coeffs(const NcoDev* __restrict__ nco, int nch, int groups, float* __restrict__ tab)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= groups * 128 * 48) return;
    const int k = idx % 48, c = idx / 48, j = 47 - k;
    double ar = 0.0, ai = 0.0;
    if (c < nch && j < 46) {
        const unsigned long long ph = 0ull - nco[c].inc * (unsigned long long)j;       // -j * inc, turns * 2^64 (wraps)
        double sn, cs;
        sincospi((double)(long long)ph * (1.0 / 9223372036854775808.0), &sn, &cs);
        ar = (double)c_cic4[j] * cs;
        ai = (double)c_cic4[j] * sn;
    }
    const float rh = tf32_rn((float)ar), rl = tf32_rn((float)(ar - (double)rh));
    const float ih = tf32_rn((float)ai), il = tf32_rn((float)(ai - (double)ih));
    float* row = tab + (size_t)c * 192;
    row[k] = rh;
    row[48 + k] = rl;
    row[96 + k] = ih;
    row[144 + k] = il;
}

// 4 consecutive samples starting at i0 (i0 % 4 == 0) for the cases off the fast path: inside the saved halo (i0 < 0),
// past the block end (zeros; only feeds outputs nobody keeps) or an integer wire format
struct TcChunk { float4 a, b; };
__device__ __noinline__ TcChunk tc_load_chunk_slow(const void* __restrict__ x, int fmt, const float2* __restrict__ halo_cur, int i0, int L)
{
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (i0 < 0) {
        const float4* h = reinterpret_cast<const float4*>(halo_cur + (kHaloMax + i0));
        a = __ldg(h);
        b = __ldg(h + 1);
    } else if (i0 >= L) {
    } else if (fmt == 0) {
        const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(x) + i0);
        a = __ldg(q);
        b = __ldg(q + 1);
    } else if (fmt == 1) {
        const int4 v = __ldg(reinterpret_cast<const int4*>(reinterpret_cast<const short2*>(x) + i0));
        a = make_float4((float)(short)(v.x & 0xffff), (float)(short)(v.x >> 16), (float)(short)(v.y & 0xffff), (float)(short)(v.y >> 16));
        b = make_float4((float)(short)(v.z & 0xffff), (float)(short)(v.z >> 16), (float)(short)(v.w & 0xffff), (float)(short)(v.w >> 16));
    } else {
        const float2 s0 = fetch_sample(x, 2, i0), s1 = fetch_sample(x, 2, i0 + 1), s2 = fetch_sample(x, 2, i0 + 2), s3 = fetch_sample(x, 2, i0 + 3);
        a = make_float4(s0.x, s0.y, s1.x, s1.y);
        b = make_float4(s2.x, s2.y, s3.x, s3.y);
    }
    TcChunk r;
    r.a = a;
    r.b = b;
    return r;
}

// fp16 form of the same table: [channel][Ar_hi, Ar_lo, Ai_hi, Ai_lo][24 words], word j = (element 2j | element 2j+1 << 16),
// which is how a 16-bit A operand sits in TMEM (one K pair per 32-bit column); the row stride stays 192 words
__global__ void k_tc_coeffs_f16(const NcoDev* __restrict__ nco, int nch, int groups, uint32_t* __restrict__ tab)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= groups * 128 * 24) return;
    const int kp = idx % 24, c = idx / 24;
    uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int k = 2 * kp + h, j = 47 - k;
        double ar = 0.0, ai = 0.0;
        if (c < nch && j < 46) {
            const unsigned long long ph = 0ull - nco[c].inc * (unsigned long long)j;
            double sn, cs;
            sincospi((double)(long long)ph * (1.0 / 9223372036854775808.0), &sn, &cs);
            ar = (double)c_cic4[j] * cs;
            ai = (double)c_cic4[j] * sn;
        }
        const __half rh = __double2half(ar), ih = __double2half(ai);
        const __half rl = __double2half(ar - (double)__half2float(rh)), il = __double2half(ai - (double)__half2float(ih));
        w[0] |= (uint32_t)__half_as_ushort(rh) << (16 * h);
        w[1] |= (uint32_t)__half_as_ushort(rl) << (16 * h);
        w[2] |= (uint32_t)__half_as_ushort(ih) << (16 * h);
        w[3] |= (uint32_t)__half_as_ushort(il) << (16 * h);
    }
    uint32_t* row = tab + (size_t)c * 192;
#pragma unroll
    for (int pl = 0; pl < 4; pl++) row[24 * pl + kp] = w[pl];
}

template <int NCR, int NHB, int G0, int G1> struct TcStrip {
    static __device__ __forceinline__ void run(const float* re, const float* im, float2& S, float2 w, CicSt* st, float2* ev, Hb11St* hs,
                                               TcEmit& em)
    {
        float2 v = cmul(make_float2(re[G0], im[G0]), S);
        S = cmul(S, w);
        cic_feed<NCR, NHB, 0, G0, TcEmit>(v, st, ev, hs, 0LL, em);
        TcStrip<NCR, NHB, G0 + 1, G1>::run(re, im, S, w, st, ev, hs, em);
    }
};
template <int NCR, int NHB, int G1> struct TcStrip<NCR, NHB, G1, G1> {
    static __device__ __forceinline__ void run(const float*, const float*, float2&, float2, CicSt*, float2*, Hb11St*, TcEmit&) {}
};

template <int NCR, int NHB, bool F16>
__global__ void __launch_bounds__(kTcThreads, 1)
    k_mix_tc(const void* __restrict__ x, int fmt, const float2* __restrict__ halo_cur, float2* __restrict__ halo_next, int L, int seg_len,
             const float* __restrict__ coef_tab, const NcoDev* __restrict__ nco,
             const unsigned long long* __restrict__ phase_cur, unsigned long long* __restrict__ phase_next, int nch, OutDesc od, float scale)
{
    typedef TcCfg<NCR, NHB> Cfg;
    // F16 = the form for int16 wire samples (fmt 1). An int16 splits EXACTLY into two fp16 numbers, v / 64 = hi + lo with
    // hi = floor((v + 32768) / 64) - 512 (an integer in [-512, 511]) and lo = ((v + 32768) mod 64) / 64, so the Hankel
    // operand is exact, all four partial products (A_hi + A_lo)(X_hi + X_lo) are kept, and kind::f16 runs at twice the
    // tf32 rate with K = 16 per MMA: 24 MMAs per accumulator and tile instead of 36, no input-dependent scale, and an
    // error that only comes from the 22-bit coefficient split. A 16-byte shared-memory chunk holds 8 samples, an X row
    // is 2 chunks = ONE MMA K-step, an A plane is 24 TMEM columns. Samples that are not int16 in this launch (the saved
    // halo of the previous block, which is float32) go through a float -> fp16 hi/lo split with the same 1/64 scale.
    constexpr int kPl = F16 ? 2 * kTcP : kTcPlane;            // bytes per B plane
    constexpr int kACols = F16 ? 24 : 48;                     // TMEM columns per A plane
    constexpr int kKSteps = F16 ? 3 : 6;
    constexpr float s_in = F16 ? 1.0f / 64.0f : 1.0f, s_out = F16 ? 64.0f : 1.0f;
    extern __shared__ __align__(128) unsigned char tc_smem[];
    unsigned char* sB = tc_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(tc_smem + kTcBarOff);
    uint64_t* b_full = bars;                                  // [kTcStages] a producer warp filled the stage
    uint64_t* b_empty = bars + kTcStages;                     // [kTcStages] the MMAs that read the stage completed
    uint64_t* acc_full = bars + 2 * kTcStages;                // [kTcSegs][2] the tile's MMAs completed (slot = tile parity)
    uint64_t* acc_empty = acc_full + 2 * kTcSegs;             // [kTcSegs][2] the epilogue set pulled the tile out of TMEM
    uint64_t* a_ready = acc_empty + 2 * kTcSegs;              // A planes are in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_ready + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int kMmaWarp = 4 * kTcSegs + 4;                 // first of the 2 kTcSegs MMA-issuing warps

    // segment e of this CTA: start sample, length, first fs/16 output computed (negative inside the halo), tile count.
    // Pure arithmetic on purpose (no per-segment arrays: dynamic indexing would put them in local memory).
    auto seg_geom = [&](int e, int& t0, int& n_seg, int& m0, int& tiles) {
        const long long t = (long long)(kTcSegs * blockIdx.x + e) * seg_len;
        t0 = (int)min(t, (long long)L);
        n_seg = max(0, min(seg_len, L - t0));
        m0 = t0 / 16 - Cfg::PRE;
        tiles = n_seg > 0 ? (n_seg / 16 + Cfg::PRE + kTcN - 1) / kTcN : 0;
    };
    auto seg_tiles = [&](int e) {
        int a, b, c2, t;
        seg_geom(e, a, b, c2, t);
        return t;
    };
    const int max_tiles = seg_tiles(0);          // segment 0 of a CTA is never shorter than its later ones

    if (warp == kMmaWarp) {
        if (lane == 0) {
            for (int i = 0; i < kTcStages; i++) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 64;" ::"r"(smem_u32(b_full + i)));      // two producer warps
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_u32(b_empty + i)));      // the re and the im issuer
            }
            for (int i = 0; i < 2 * kTcSegs; i++) {
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 2;" ::"r"(smem_u32(acc_full + i)));
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(acc_empty + i)));
            }
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(a_ready)), "r"(128 * kTcSegs));
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    if (blockIdx.x == gridDim.x - 1 && blockIdx.y == 0) save_halo(x, fmt, halo_cur, halo_next, L);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *tmem_slot;

    if (warp >= 4 * kTcSegs && warp < kMmaWarp) {
        // ===== producers: wideband samples -> tf32 hi/lo planes of the Hankel operand. Warp pair (pw >> 1) takes every
        // other tile, each warp of the pair stages half of the tile's 16-byte chunks; two tiles are in flight. =====
        const int pw = warp - 4 * kTcSegs;
        constexpr int kChunks = 4 * kTcRows;                              // 16-byte chunks (4 samples) per plane and tile
        constexpr int kHalf = (kChunks + 1) / 2;
        constexpr int kRounds = (kHalf + 31) / 32;
        const int c_lo = (pw & 1) * kHalf, c_hi = min(kChunks, c_lo + kHalf);
        // The global loads of a tile are issued one tile ahead of its conversion (DRAM latency is ~1 us, more than a
        // tile time): two register sets alternate.
        struct Pending { int st, my; bool raw; };
        // returns true when the registers hold RAW int16 samples (F16 form, whole tile inside the block): va[j] then
        // carries four (I, Q) words and vb[j] is unused
        auto issue_loads = [&](int mt, float4* va, float4* vb) -> bool {
            const int i_lo = 16 * (mt - 2);
            const bool inside = i_lo >= 0 && i_lo + 16 * kTcRows <= L;
            const bool fast = fmt == 0 && inside;
            const bool raw = F16 && fmt == 1 && inside;
#pragma unroll
            for (int j = 0; j < kRounds; j++) {
                const int idx = c_lo + lane + 32 * j;                     // chunk (row idx / 4, q = idx % 4)
                const int i0 = i_lo + 4 * idx;                            // first of 4 consecutive samples; i0 % 4 == 0
                va[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                vb[j] = va[j];
                if (idx < c_hi) {
                    if (raw) {
                        va[j] = __ldg(reinterpret_cast<const float4*>(reinterpret_cast<const short2*>(x) + i0));
                    } else if (fast) {
                        const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(x) + i0);
                        va[j] = __ldg(q);
                        vb[j] = __ldg(q + 1);
                    } else {
                        const TcChunk ch = tc_load_chunk_slow(x, fmt, halo_cur, i0, L);
                        va[j] = ch.a;
                        vb[j] = ch.b;
                    }
                }
            }
            return raw;
        };
        auto flush = [&](const Pending& pd, const float4* va, const float4* vb) {
            // convert a tile's samples (in registers) and hand the stage to the MMA warps
            unsigned char* pl = sB + pd.st * kTcStage;
            bar_wait(b_empty + pd.st, ((pd.my / kTcStages) & 1) ^ 1);
#pragma unroll
            for (int j = 0; j < kRounds; j++) {
                const int idx = c_lo + lane + 32 * j;
                if (idx < c_hi) {
                    const float4 a = va[j], b = vb[j];
                    if constexpr (F16) {
                        // 4 samples = half of a 16-byte chunk: row idx / 4, chunk (idx / 2) % 2, half idx % 2
                        const int off = 16 * (idx >> 2) + kTcP * ((idx >> 1) & 1) + 8 * (idx & 1);
                        if (pd.raw) {
                            // (I, Q) int16 words of samples 0..3 -> (I0 I1), (I2 I3), (Q0 Q1), (Q2 Q3); per pair: u = v + 32768,
                            // fp16 bit patterns 0x6400 | field are 1024 + field, so hi = (1024 + u / 64) - 1536 and
                            // lo = (1024 + u % 64) / 64 - 16, both exact
                            const uint32_t w0 = __float_as_uint(a.x), w1 = __float_as_uint(a.y), w2 = __float_as_uint(a.z), w3 = __float_as_uint(a.w);
                            const uint32_t pk[4] = {__byte_perm(w0, w1, 0x5410), __byte_perm(w2, w3, 0x5410), __byte_perm(w0, w1, 0x7632),
                                                    __byte_perm(w2, w3, 0x7632)};
                            uint32_t hi[4], lo[4];
                            const __half2 c1536 = __floats2half2_rn(1536.f, 1536.f), c64 = __floats2half2_rn(1.f / 64.f, 1.f / 64.f),
                                          c16 = __floats2half2_rn(-16.f, -16.f);
#pragma unroll
                            for (int h = 0; h < 4; h++) {
                                const uint32_t u = pk[h] ^ 0x80008000u;
                                const uint32_t hb = ((u >> 6) & 0x03ff03ffu) | 0x64006400u, lb = (u & 0x003f003fu) | 0x64006400u;
                                const __half2 hv = __hsub2(*reinterpret_cast<const __half2*>(&hb), c1536);
                                const __half2 lv = __hfma2(*reinterpret_cast<const __half2*>(&lb), c64, c16);
                                hi[h] = *reinterpret_cast<const uint32_t*>(&hv);
                                lo[h] = *reinterpret_cast<const uint32_t*>(&lv);
                            }
                            *reinterpret_cast<uint2*>(pl + off) = make_uint2(hi[0], hi[1]);
                            *reinterpret_cast<uint2*>(pl + kPl + off) = make_uint2(lo[0], lo[1]);
                            *reinterpret_cast<uint2*>(pl + 2 * kPl + off) = make_uint2(hi[2], hi[3]);
                            *reinterpret_cast<uint2*>(pl + 3 * kPl + off) = make_uint2(lo[2], lo[3]);
                            continue;
                        }
                        // samples that did not arrive as whole int16 tiles (the float32 halo of the previous block, tiles that
                        // straddle the block ends): the SAME split in float arithmetic -- u = x + 32768, hi = floor(u / 64) - 512,
                        // lo = (u mod 64) / 64 -- so a sample's operand never depends on which tile or segment carried it
                        const float xr[4] = {a.x, a.z, b.x, b.z};
                        const float xi[4] = {a.y, a.w, b.y, b.w};
                        __half2 rh[2], rl[2], ih[2], il[2];
                        auto split = [&](float x, float& hi, float& lo) {
                            const float u = x + 32768.0f;
                            const float q = floorf(u * s_in);
                            hi = q - 512.0f;
                            lo = fmaf(q, -64.0f, u) * s_in;
                        };
#pragma unroll
                        for (int h = 0; h < 2; h++) {
                            float h0, l0, h1, l1;
                            split(xr[2 * h], h0, l0);
                            split(xr[2 * h + 1], h1, l1);
                            rh[h] = __floats2half2_rn(h0, h1);
                            rl[h] = __floats2half2_rn(l0, l1);
                            split(xi[2 * h], h0, l0);
                            split(xi[2 * h + 1], h1, l1);
                            ih[h] = __floats2half2_rn(h0, h1);
                            il[h] = __floats2half2_rn(l0, l1);
                        }
                        auto put = [&](int plane, const __half2* v) {
                            uint2 u;
                            u.x = *reinterpret_cast<const uint32_t*>(&v[0]);
                            u.y = *reinterpret_cast<const uint32_t*>(&v[1]);
                            *reinterpret_cast<uint2*>(pl + plane * kPl + off) = u;
                        };
                        put(0, rh); put(1, rl); put(2, ih); put(3, il);
                    } else {
                        const float4 rh = make_float4(tf32_rn(a.x), tf32_rn(a.z), tf32_rn(b.x), tf32_rn(b.z));
                        const float4 ih = make_float4(tf32_rn(a.y), tf32_rn(a.w), tf32_rn(b.y), tf32_rn(b.w));
                        const float4 rl = make_float4(tf32_rn(a.x - rh.x), tf32_rn(a.z - rh.y), tf32_rn(b.x - rh.z), tf32_rn(b.z - rh.w));
                        const float4 il = make_float4(tf32_rn(a.y - ih.x), tf32_rn(a.w - ih.y), tf32_rn(b.y - ih.z), tf32_rn(b.w - ih.w));
                        const int off = 16 * (idx >> 2) + kTcP * (idx & 3);
                        *reinterpret_cast<float4*>(pl + off) = rh;
                        *reinterpret_cast<float4*>(pl + kPl + off) = rl;
                        *reinterpret_cast<float4*>(pl + 2 * kPl + off) = ih;
                        *reinterpret_cast<float4*>(pl + 3 * kPl + off) = il;
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            bar_arrive(b_full + pd.st);
        };
        // this warp pair's tiles are those of segment (pw >> 1): tile k of the segment is global tile number 2 k + e
        // while both segments are alive, so the (stage, use count) sequence is recomputed the same way every role does
        float4 va0[kRounds], vb0[kRounds], va1[kRounds], vb1[kRounds];
        Pending p0 = {-1, 0, false}, p1 = {-1, 0, false};
        int cnt = 0, mine = 0;
        for (int k = 0; k < max_tiles; k++) {
#pragma unroll 1
            for (int e = 0; e < kTcSegs; e++) {
                int t0, n_seg, m0, tiles;
                seg_geom(e, t0, n_seg, m0, tiles);
                if (k >= tiles) continue;
                const int my = cnt++;
                if ((my & 1) != (pw >> 1)) continue;
                Pending now = {my % kTcStages, my, false};
                if ((mine++ & 1) == 0) {
                    now.raw = issue_loads(m0 + kTcN * k, va0, vb0);
                    p0 = now;
                    if (p1.st >= 0) { flush(p1, va1, vb1); p1.st = -1; }
                } else {
                    now.raw = issue_loads(m0 + kTcN * k, va1, vb1);
                    p1 = now;
                    if (p0.st >= 0) { flush(p0, va0, vb0); p0.st = -1; }
                }
            }
        }
        // drain in issue order
        if ((mine & 1) == 0) { if (p0.st >= 0) flush(p0, va0, vb0); if (p1.st >= 0) flush(p1, va1, vb1); }
        else { if (p1.st >= 0) flush(p1, va1, vb1); if (p0.st >= 0) flush(p0, va0, vb0); }
    } else if (warp >= kMmaWarp) {
        // ===== MMA issuers: one warp per (segment, re | im accumulator); every accumulator has exactly one writer =====
        const int me = (warp - kMmaWarp) >> 1, half = (warp - kMmaWarp) & 1;
        // instruction descriptor: D = fp32; A/B format tf32 (2) or fp16 (0); N, M
        const uint32_t idesc = (1u << 4) | (F16 ? 0u : ((2u << 7) | (2u << 10))) | ((uint32_t)(kTcN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const uint32_t idesc_na = idesc | (1u << 13);       // negate A
        const uint32_t b_hi32 = (128u >> 4) | (1u << 14);   // SBO = 128 B (8 rows of 16 B), descriptor version 1
        const uint32_t lo_lbo = ((uint32_t)kTcP >> 4) << 16;
        const uint32_t d0 = tm + (uint32_t)(kTcAccCol + 4 * kTcN * me + kTcN * half);
        bar_wait(a_ready, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        int cnt = 0;
        for (int k = 0; k < max_tiles; k++) {
#pragma unroll 1
            for (int e = 0; e < kTcSegs; e++) {
                if (k >= seg_tiles(e)) continue;
                const int my = cnt++;
                const int slot = k & 1;
                if (e != me) continue;
                const int st = my % kTcStages;
                bar_wait(b_full + st, (my / kTcStages) & 1);
                bar_wait(acc_empty + 2 * e + slot, ((k >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t b0 = smem_u32(sB + st * kTcStage);
                    const uint32_t d = d0 + (uint32_t)(2 * kTcN * slot);
                    // planes: A 0=Ar_hi 1=Ar_lo 2=Ai_hi 3=Ai_lo (48 TMEM columns each);  B 0=Xr_hi 1=Xr_lo 2=Xi_hi 3=Xi_lo.
                    // D_re = Ar Xr - Ai Xi,  D_im = Ar Xi + Ai Xr; small terms (lo*hi, hi*lo) first.
                    // tf32: lo*lo is dropped (2^-22 relative); fp16 (exact int16 operand): all four partial products
                    constexpr int kTerms = F16 ? 8 : 6;
                    constexpr int a_pl[8] = {1, 0, 3, 2, 0, 2, 1, 3};
                    constexpr int b_re[8] = {0, 1, 2, 3, 0, 2, 1, 3}, n_re[8] = {0, 0, 1, 1, 0, 1, 0, 1};
                    constexpr int b_im[8] = {2, 3, 0, 1, 2, 0, 3, 1};
                    // issue order: the two lo*lo terms (fp16 only) first, then as before
                    constexpr int order[8] = {6, 7, 0, 1, 2, 3, 4, 5};
#pragma unroll
                    for (int tt = 0; tt < 8; tt++) {
                        const int t = order[tt];
                        if (t >= kTerms) continue;
                        const bool first_term = F16 ? tt == 0 : tt == 2;
                        const int ap = a_pl[t];
                        const int bp = half ? b_im[t] : b_re[t];
                        const uint32_t id = (!half && n_re[t]) ? idesc_na : idesc;
#pragma unroll
                        for (int ks = 0; ks < kKSteps; ks++) {
                            // tf32: K-step = 8 samples = chunks 2 ks, 2 ks + 1 of the row pair; fp16: K-step = one whole row
                            const uint32_t baddr = F16 ? b0 + bp * kPl + 16 * ks : b0 + bp * kPl + 16 * (ks >> 1) + kTcP * ((2 * ks) & 3);
                            const uint64_t bd = ((uint64_t)b_hi32 << 32) | (uint64_t)(((baddr >> 4) & 0x3fff) | lo_lbo);
                            const uint32_t acc = (first_term && ks == 0) ? 0u : 1u;
                            if constexpr (F16) tc_mma_ts_f16(d, tm + (uint32_t)(kACols * ap + 8 * ks), bd, id, acc);
                            else tc_mma_ts(d, tm + (uint32_t)(kACols * ap + 8 * ks), bd, id, acc);
                        }
                    }
                    tc_commit(b_empty + st);
                    tc_commit(acc_full + 2 * e + slot);
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue set e (warps 4e .. 4e+3): TMEM lane = channel =====
        const int e = warp >> 2;
        const int ltid = tid & 127;
        const int c = blockIdx.y * 128 + ltid;
        const bool valid = c < nch;
        const uint32_t lane_base = tm + ((uint32_t)((warp & 3) * 32) << 16);
        {
            // set e puts A planes 2e, 2e+1 of its 128 rows into TMEM columns 2 kACols e .. 2 kACols (e + 1) - 1
            const float4* row = reinterpret_cast<const float4*>(coef_tab + (size_t)c * 192 + 2 * kACols * e);
#pragma unroll
            for (int blk = 0; blk < 2 * kACols / 16; blk++) {
                float v[16];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float4 f = __ldg(row + 4 * blk + i);
                    v[4 * i] = f.x; v[4 * i + 1] = f.y; v[4 * i + 2] = f.z; v[4 * i + 3] = f.w;
                }
                tmem_st16(lane_base + (uint32_t)(2 * kACols * e + 16 * blk), v);
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(a_ready);
        }
        int t0, n_seg, m_start, ntiles;
        seg_geom(e, t0, n_seg, m_start, ntiles);
        NcoDev p = nco[valid ? c : 0];
        const unsigned long long ph0 = phase_cur[valid ? c : 0];
        if (blockIdx.x == 0 && e == 0 && valid) phase_next[c] = ph0 + (unsigned long long)L * p.inc;
        const float2 w16 = make_float2(p.wtc, p.wts);
        CicSt st[NCR > 0 ? NCR : 1];
        float2 ev[NCR > 0 ? NCR : 1];
        Hb11St hs[NHB > 0 ? NHB : 1];
#pragma unroll
        for (int i = 0; i < (NCR > 0 ? NCR : 1); i++) {
            st[i].xodd = make_float2(0.f, 0.f);
            st[i].xeven = make_float2(0.f, 0.f);
            ev[i] = make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int i = 0; i < (NHB > 0 ? NHB : 1); i++) {
#pragma unroll
            for (int k = 0; k < 5; k++) hs[i].e[k] = make_float2(0.f, 0.f);
#pragma unroll
            for (int k = 0; k < 3; k++) hs[i].o[k] = make_float2(0.f, 0.f);
        }
        constexpr int SH = NCR + NHB;                                   // fs/16 outputs per emitted row = 2^SH
        const long long row_lo = (long long)(t0 / (16 << SH));
        const long long q_first = (long long)(m_start >> SH);            // exact: 2^SH divides PRE and t0/16
        TcEmit em;
        em.mask = od.transposed ? (unsigned)(kDecRing - 1) : od.mask;
        em.estride = od.transposed ? 1 : od.stride;
        em.p0 = od.transposed ? od.p + (size_t)c * kDecRing : od.p + c;
        em.pos = (unsigned)((od.base + q_first) & (long long)em.mask);
        em.rel = (int)(q_first - row_lo);
        em.n_rows = valid ? (int)((long long)((t0 + n_seg) / (16 << SH)) - row_lo) : 0;
        em.scale = scale * s_out;
        em.h0 = c_hb_taps[0]; em.h2 = c_hb_taps[1]; em.h4 = c_hb_taps[2];
        // Oscillator at fs/16: output m is rotated by the phase of its newest input sample 16m+15, P + (16m+16) inc.
        // It is re-seeded exactly (sincospi of the 64-bit phase) at ABSOLUTE multiples of 32 outputs and rotated in
        // float32 in between; a segment that starts between two seed points replays the rotations since the last one,
        // so every output bit is independent of how the block was cut into segments and tiles.
        float2 S;
        {
            const int lead = ((m_start % 32) + 32) % 32;                 // 0 or 16
            S = seed_osc(ph0 + (unsigned long long)(long long)(16 * (m_start - lead) + 16) * p.inc);
            for (int i = 0; i < lead; i++) S = cmul(S, w16);
        }
        // Software pipeline over 16-output groups (two per tile): the tcgen05.ld of group g+1 -- and the wait for its
        // tile's MMAs -- is in flight while group g runs through the oscillator / CIC / half-band arithmetic. The
        // accumulator slot goes back to the MMA warps as soon as the tile's second group is in registers.
        static_assert(kTcN == 32, "the epilogue pipeline below assumes two 16-output groups per tile");
        float reA[16], imA[16], reB[16], imB[16];
        auto acc_of = [&](int k) { return lane_base + (uint32_t)(kTcAccCol + 4 * kTcN * e + 2 * kTcN * (k & 1)); };
        auto group_math = [&](const float* re, const float* im, int m_first) {
            if ((m_first & 31) == 0) S = seed_osc(ph0 + (unsigned long long)(long long)(16 * m_first + 16) * p.inc);
            TcStrip<NCR, NHB, 0, 16>::run(re, im, S, w16, st, ev, hs, em);
        };
        if (ntiles > 0) {
            bar_wait(acc_full + 2 * e, 0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            tmem_ld16(acc_of(0), reA);
            tmem_ld16(acc_of(0) + kTcN, imA);
        }
        for (int k = 0; k < ntiles; k++) {
            const int slot = k & 1;
            const int mt = m_start + kTcN * k;
            // group 0 of tile k is in flight (A registers): complete it, start group 1 (B registers)
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tmem_pin16(reA);
            tmem_pin16(imA);
            tmem_ld16(acc_of(k) + 16, reB);
            tmem_ld16(acc_of(k) + kTcN + 16, imB);
            group_math(reA, imA, mt);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tmem_pin16(reB);
            tmem_pin16(imB);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            bar_arrive(acc_empty + 2 * e + slot);
            if (k + 1 < ntiles) {
                bar_wait(acc_full + 2 * e + ((k + 1) & 1), ((k + 1) >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                tmem_ld16(acc_of(k + 1), reA);
                tmem_ld16(acc_of(k + 1) + kTcN, imA);
            }
            group_math(reB, imB, mt + 16);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kMmaWarp) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(512));
}

typedef void (*K1TFn)(const void*, int, const float2*, float2*, int, int, const float*, const NcoDev*,
                      const unsigned long long*, unsigned long long*, int, OutDesc, float);
static K1TFn k1t_kernel(int ncr, int nhb, bool f16)
{
    static const K1TFn t32[3][4] = {{k_mix_tc<0, 0, false>, k_mix_tc<0, 1, false>, k_mix_tc<0, 2, false>, k_mix_tc<0, 3, false>},
                                    {k_mix_tc<1, 0, false>, k_mix_tc<1, 1, false>, k_mix_tc<1, 2, false>, k_mix_tc<1, 3, false>},
                                    {k_mix_tc<2, 0, false>, k_mix_tc<2, 1, false>, k_mix_tc<2, 2, false>, k_mix_tc<2, 3, false>}};
    static const K1TFn t16[3][4] = {{k_mix_tc<0, 0, true>, k_mix_tc<0, 1, true>, k_mix_tc<0, 2, true>, k_mix_tc<0, 3, true>},
                                    {k_mix_tc<1, 0, true>, k_mix_tc<1, 1, true>, k_mix_tc<1, 2, true>, k_mix_tc<1, 3, true>},
                                    {k_mix_tc<2, 0, true>, k_mix_tc<2, 1, true>, k_mix_tc<2, 2, true>, k_mix_tc<2, 3, true>}};
    return f16 ? t16[ncr][nhb] : t32[ncr][nhb];
}

// ------------------------------------------------------------------------------------------
// K2a: NS consecutive 11-tap half-band stages in ONE pass over HBM (the stages right after
// kernel 1 carry most of kernel 2's traffic). A CTA owns CH (16) channels x T final outputs: it stages
// the 2^NS*T + halo input rows in shared memory (all 8 warps issue whole-row 256-byte loads, ~30 in
// flight per warp), then runs the stages time-parallel out of shared memory (lane = channel, so a
// warp reads one row per LDS.64, conflict-free), each stage writing the next stage's rows, the last
// one writing HBM. Only the first stage's input and the last stage's output touch HBM.
// Row counts: c[NS] = T, c[j] = 2 c[j+1] + 9 (an output needs inputs 2m-10 .. 2m).
// ------------------------------------------------------------------------------------------
constexpr int kHbcCh = 16;        // channels per CTA of the fused half-band kernels
constexpr int kHbc2T = 96;        // final outputs per CTA, two fused stages
constexpr int kHbc3T = 48;        // three fused stages
template <int NS, int T> struct HbcCfg {
    static constexpr int c(int j) { return j >= NS ? T : 2 * c(j + 1) + 9; }
    static constexpr int smem_rows() { return NS == 2 ? c(0) + c(1) : c(0) + c(1) + c(2); }
};

template <int NS, int T, int CH>
__global__ void __launch_bounds__(256) k_hb11_chain(const float2* __restrict__ in, unsigned in_mask, long long in_base, int stride,
                                                    int n_out, OutDesc od)
{
    // CH channels per CTA (16: a row is one 128-byte line, which halves shared memory per output and
    // lets T grow, i.e. less halo); a warp covers 32/CH rows per instruction
    typedef HbcCfg<NS, T> Cfg;
    constexpr int RPW = 32 / CH;                   // rows per warp instruction
    extern __shared__ float2 sm_rows[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / CH, l = lane % CH;
    const int c = blockIdx.y * CH + l;
    const int o0 = blockIdx.x * T;
    const float h0 = c_hb_taps[0], h2 = c_hb_taps[1], h4 = c_hb_taps[2];
    // first input row of every stage for this tile: lo[NS] = o0, lo[j] = 2 lo[j+1] - 10
    long long lo0 = o0;
#pragma unroll
    for (int j = 0; j < NS; j++) lo0 = 2 * lo0 - 10;
    float2* buf = sm_rows;
    {
        // stage the input rows with cp.async (LDGSTS): 16 bytes = 2 channels per lane, no registers
        // held -- the whole tile is in flight at once
        constexpr int LPR = CH / 2;                // lanes per row
        constexpr int RPI = 32 / LPR;              // rows per warp instruction
        const int rsub = lane / LPR, l2 = lane % LPR;
        const int cbase = blockIdx.y * CH + 2 * l2;
        for (int r = RPI * warp + rsub; r < Cfg::c(0); r += 8 * RPI) {
            const float2* g = in + (size_t)((in_base + lo0 + r) & in_mask) * stride + cbase;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + r * CH + 2 * l2);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const int n_j = Cfg::c(j + 1);             // outputs of this stage
        float2* nxt = buf + Cfg::c(j) * CH;
        for (int i = RPW * warp + sub; i < n_j; i += 8 * RPW) {
            const float2* x = buf + (2 * i) * CH + l;
            const float2 a0 = x[0], a2 = x[2 * CH], a4 = x[4 * CH], a5 = x[5 * CH], a6 = x[6 * CH], a8 = x[8 * CH], a10 = x[10 * CH];
            float2 y;
            // same operation order as k_halfband, so a stage gives identical bits on either path
            y.x = fmaf(h4, a4.x + a6.x, fmaf(h2, a2.x + a8.x, fmaf(h0, a0.x + a10.x, 0.5f * a5.x)));
            y.y = fmaf(h4, a4.y + a6.y, fmaf(h2, a2.y + a8.y, fmaf(h0, a0.y + a10.y, 0.5f * a5.y)));
            if (j + 1 < NS) nxt[i * CH + l] = y;
            else if (o0 + i < n_out) store_out(od, o0 + i, c, y);
        }
        if (j + 1 < NS) __syncthreads();
        buf = nxt;
    }
}

// ------------------------------------------------------------------------------------------
// K2b: the LAST stages of the ladder (any half-band lengths, up to 4) in one pass. Their data is small (a few
// thousand rows per block) and as separate launches they are launch- and tail-latency bound; here a CTA owns CH
// channels x T final outputs, stages c[0] = 2 c[1] + len[0] - 2 input rows in shared memory and runs the stages
// out of it, exactly like K2a. Same operation order as k_halfband (centre tap, then the outermost pair inwards),
// so either path gives the same bits.
// ------------------------------------------------------------------------------------------
constexpr int kTailMaxPerThread = 12;     // outputs of one stage a thread of k_hb_tail holds in registers
struct TailCfg {
    int ns;
    int len[4];
    int toff[4];
    int c[5];          // rows: c[ns] = T, c[j] = 2 c[j+1] + len[j] - 2
};

template <int CH>
__global__ void __launch_bounds__(256, 4) k_hb_tail(const float2* __restrict__ in, unsigned in_mask, long long in_base, int stride, int n_out,
                                                 TailCfg cfg, OutDesc od)
{
    constexpr int RPW = 32 / CH;
    extern __shared__ float2 sm_rows[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int sub = lane / CH, l = lane % CH;
    const int c = blockIdx.y * CH + l;
    const int o0 = blockIdx.x * cfg.c[cfg.ns];
    long long lo0 = o0;
    for (int j = cfg.ns - 1; j >= 0; j--) lo0 = 2 * lo0 - (cfg.len[j] - 1);
    float2* buf = sm_rows;
    {
        constexpr int LPR = CH / 2;                // lanes per row (16 bytes = 2 channels per lane)
        constexpr int RPI = 32 / LPR;
        const int rsub = lane / LPR, l2 = lane % LPR;
        const int cbase = blockIdx.y * CH + 2 * l2;
        for (int r = RPI * warp + rsub; r < cfg.c[0]; r += 8 * RPI) {
            const float2* g = in + (size_t)((in_base + lo0 + r) & in_mask) * stride + cbase;
            const unsigned sa = (unsigned)__cvta_generic_to_shared(buf + r * CH + 2 * l2);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();
    // A stage's outputs stay in registers until every thread has read its inputs, then overwrite the front of the SAME
    // buffer: the CTA needs c[0] rows of shared memory instead of c[0] + c[1] + ..., which lets the tile be 1.5x longer at
    // the same number of resident CTAs -- the whole grid in ONE wave (1024 CTAs on 740 slots ran as two).
    for (int j = 0; j < cfg.ns; j++) {
        const int N = cfg.len[j], half = (N - 1) >> 1, n_j = cfg.c[j + 1];
        const bool last = j + 1 == cfg.ns;
        float2 outv[kTailMaxPerThread];
#pragma unroll
        for (int q = 0; q < kTailMaxPerThread; q++) {
            const int i = RPW * warp + sub + q * 8 * RPW;
            float2 acc = make_float2(0.f, 0.f);
            if (i < n_j) {
                const float2* x = buf + (2 * i) * CH + l;          // oldest tap of output i
                const float2 xc = x[half * CH];
                acc = make_float2(0.5f * xc.x, 0.5f * xc.y);
                int t = cfg.toff[j];
                for (int k = 0; k < half; k += 2, t++) {
                    const float h = c_hb_taps[t];
                    const float2 u = x[k * CH], v = x[(N - 1 - k) * CH];
                    acc.x = fmaf(h, u.x + v.x, acc.x);
                    acc.y = fmaf(h, u.y + v.y, acc.y);
                }
                if (last && o0 + i < n_out) store_out(od, o0 + i, c, acc);
            }
            outv[q] = acc;
        }
        if (!last) {
            __syncthreads();
#pragma unroll
            for (int q = 0; q < kTailMaxPerThread; q++) {
                const int i = RPW * warp + sub + q * 8 * RPW;
                if (i < n_j) buf[i * CH + l] = outv[q];
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// K2s: the first THREE half-band stages after kernel 1 (11, 11 and 11 or 15 taps: 7/8 of kernel 2's input bytes) in one
// streaming pass, state in REGISTERS. Lane = channel (32 channels = one 256-byte row of the time-major stage ring),
// warp = time slice: every warp walks its slice row by row through all three stages with register delay lines, like
// kernel 1's epilogue does for its fused stages, so a row costs ONE shared-memory load and ~7 packed FP32x2
// instructions per channel instead of a gather of 7-8 loads per output and stage -- the pass is bound by HBM, not by
// instruction issue. Rows arrive through a per-warp ring of TMA boxes (cp.async.bulk.tensor.2d, 32 rows x 256 B,
// completion on an mbarrier, three boxes in flight ahead of the arithmetic); warps never synchronise with each other.
// A slice re-primes its (feed-forward) delay lines from 86-117 rows in front of it: the stage ring in HBM keeps the
// previous block. Per output the operation sequence is k_halfband's (centre tap, then the outermost pair inwards;
// FADD2/FFMA2 on the (re, im) pair round exactly like the scalar pair), so every path gives the same bits.
// ------------------------------------------------------------------------------------------
constexpr int kH3Ch = 32;          // channels per warp (float2 each: 256-byte rows)
constexpr int kH3Box = 32;         // rows per TMA box
constexpr int kH3Boxes = 4;        // boxes in a warp's shared-memory ring
constexpr int kH3Warps = 2;        // time slices per CTA
constexpr int kH3Smem = kH3Warps * kH3Boxes * kH3Box * kH3Ch * 8;

template <int N> struct HbR {      // streaming half-band decimator, N = 11 or 15
    static constexpr int H = (N - 1) / 2, NP = (H + 1) / 2;      // even-sample history, tap pairs ( = odd-sample history)
    float2 e[H];                   // e[0] = most recent even-indexed input
    float2 o[NP];                  // o[0] = most recent odd-indexed input
    __device__ __forceinline__ void clear()
    {
#pragma unroll
        for (int i = 0; i < H; i++) e[i] = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NP; i++) o[i] = make_float2(0.f, 0.f);
    }
    __device__ __forceinline__ void odd(float2 v)
    {
#pragma unroll
        for (int i = NP - 1; i > 0; i--) o[i] = o[i - 1];
        o[0] = v;
    }
    // even-indexed input x[2m] completes output m
    __device__ __forceinline__ float2 even(float2 v, const float* h)
    {
        float2 acc = __fmul2_rn(make_float2(0.5f, 0.5f), o[NP - 1]);
        acc = __ffma2_rn(make_float2(h[0], h[0]), __fadd2_rn(e[H - 1], v), acc);
#pragma unroll
        for (int k = 1; k < NP; k++) acc = __ffma2_rn(make_float2(h[k], h[k]), __fadd2_rn(e[H - 1 - k], e[k - 1]), acc);
#pragma unroll
        for (int i = H - 1; i > 0; i--) e[i] = e[i - 1];
        e[0] = v;
        return acc;
    }
};

template <int L2>
__global__ void __launch_bounds__(32 * kH3Warps) k_hb3r(const __grid_constant__ CUtensorMap tmap, unsigned in_mask, long long in_base, int n_out,
                                                         int per_slice, int toff2, OutDesc od)
{
    extern __shared__ __align__(128) unsigned char h3_smem[];
    __shared__ __align__(8) uint64_t full[kH3Warps][kH3Boxes];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slice = blockIdx.x * kH3Warps + warp;
    const int o0 = slice * per_slice;
    const int o1 = min(o0 + per_slice, n_out);
    if (o0 >= o1) return;                                     // whole warp; warps are independent
    const float2* ring = reinterpret_cast<const float2*>(h3_smem) + (size_t)warp * kH3Boxes * kH3Box * kH3Ch;
    constexpr int HALO = 2 * (2 * (L2 - 1) + 10) + 10;        // stage-0 rows in front of output o0's first stage-2 input
    // the slice's stream starts on a 32-row boundary of the ring (in_base is a multiple of 32: no box straddles the ring
    // end, and a row's parity at every stage is its position in the box)
    const long long abs0 = (in_base + 8LL * o0 - HALO) & ~(long long)(kH3Box - 1);
    const int first0 = (int)(abs0 - in_base);
    const int n_boxes = (8 * (o1 - 1) - first0) / kH3Box + 1;
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kH3Boxes; i++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[warp][i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    auto issue = [&](int b) {
        const uint32_t bar = smem_u32(&full[warp][b & (kH3Boxes - 1)]);
        const int y = (int)((abs0 + (long long)b * kH3Box) & (long long)in_mask);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(kH3Box * kH3Ch * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
                         smem_u32(ring + (size_t)(b & (kH3Boxes - 1)) * kH3Box * kH3Ch)),
                     "l"(&tmap), "r"((int)(blockIdx.y * kH3Ch * 2)), "r"(y), "r"(bar)
                     : "memory");
    };
    if (lane == 0)
        for (int b = 0; b < kH3Boxes - 1 && b < n_boxes; b++) issue(b);
    float h0[3], h2[HbR<L2>::NP];
#pragma unroll
    for (int i = 0; i < 3; i++) h0[i] = c_hb_taps[i];         // 11-tap stage: table offset 0
#pragma unroll
    for (int i = 0; i < HbR<L2>::NP; i++) h2[i] = c_hb_taps[toff2 + i];
    HbR<11> s0, s1;
    HbR<L2> s2;
    s0.clear(); s1.clear(); s2.clear();
    const int c = blockIdx.y * kH3Ch + lane;
    for (int b = 0; b < n_boxes; b++) {
        // boxes 0 .. kH3Boxes-2 were issued up front; box b + kH3Boxes - 1 goes into the slot of box b - 1, which the whole
        // warp finished with one iteration ago (for b = 0 that slot has never been used): three boxes stay in flight
        if (lane == 0 && b + kH3Boxes - 1 < n_boxes) issue(b + kH3Boxes - 1);
        {
            const uint32_t bar = smem_u32(&full[warp][b & (kH3Boxes - 1)]);
            const uint32_t parity = (uint32_t)((b / kH3Boxes) & 1);
            asm volatile(
                "{\n\t.reg .pred p;\n"
                "W_%=:\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                "@p bra D_%=;\n\t"
                "bra W_%=;\n"
                "D_%=:\n\t}\n" ::"r"(bar),
                "r"(parity)
                : "memory");
        }
        const float2* x = ring + (size_t)(b & (kH3Boxes - 1)) * kH3Box * kH3Ch + lane;
        const int m_box = (first0 + b * kH3Box) >> 3;         // stage-2 output index completed by row 0 of this box
#pragma unroll
        for (int r = 0; r < kH3Box; r++) {
            const float2 v = x[r * kH3Ch];
            if (r & 1) s0.odd(v);
            else {
                const float2 y0 = s0.even(v, h0);
                if (r & 2) s1.odd(y0);
                else {
                    const float2 y1 = s1.even(y0, h0);
                    if (r & 4) s2.odd(y1);
                    else {
                        const float2 y2 = s2.even(y1, h2);
                        const int m = m_box + (r >> 3);
                        if (m >= o0 && m < o1) store_out(od, m, c, y2);
                    }
                }
            }
        }
        __syncwarp();          // every lane is done with this box before lane 0 lets the next iteration refill its neighbour
    }
}

// ------------------------------------------------------------------------------------------
// K2: one decimate-by-2 stage, thread per (output row m, channel c)
//   half-band : y[m] = sum_j h[j] x[2m-(N-1)+j]            (dsp/downconvert.cpp:286-320, 348-423)
//   N == 3    : CIC3, y[m] = .125 (x[2m+1] + 3x[2m] + 3x[2m-1] + x[2m-2])          (:444-460)
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_halfband(const float2* __restrict__ in, unsigned in_mask, long long in_base,
                                                  int stride, int n_out, int N, int tap_off, OutDesc od)
{
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int c = (int)(idx % stride);
    const long long m = idx / stride;
    if (m >= n_out) return;
    float2 acc;
    if (N == 3) {
        const long long a = in_base + 2 * m;
        const float2 x1 = in[(size_t)((a + 1) & in_mask) * stride + c];
        const float2 x0 = in[(size_t)(a & in_mask) * stride + c];
        const float2 xm1 = in[(size_t)((a - 1) & in_mask) * stride + c];
        const float2 xm2 = in[(size_t)((a - 2) & in_mask) * stride + c];
        acc.x = .125f * ((x1.x + xm2.x) + 3.0f * (xm1.x + x0.x));
        acc.y = .125f * ((x1.y + xm2.y) + 3.0f * (xm1.y + x0.y));
    } else {
        const long long a = in_base + 2 * m - (N - 1);     // oldest tap
        const int half = (N - 1) >> 1;
        const float2 xc = in[(size_t)((a + half) & in_mask) * stride + c];
        acc.x = 0.5f * xc.x;
        acc.y = 0.5f * xc.y;
        int t = tap_off;
        for (int j = 0; j < half; j += 2, t++) {
            const float h = c_hb_taps[t];
            const float2 u = in[(size_t)((a + j) & in_mask) * stride + c];
            const float2 v = in[(size_t)((a + (N - 1) - j) & in_mask) * stride + c];
            acc.x = fmaf(h, u.x + v.x, acc.x);
            acc.y = fmaf(h, u.y + v.y, acc.y);
        }
    }
    store_out(od, m, c, acc);
}

// ------------------------------------------------------------------------------------------
// NCO start-up amplitude
// ------------------------------------------------------------------------------------------
__global__ void k_unpack(const void* __restrict__ raw, int fmt, float2* __restrict__ out, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fetch_sample(raw, fmt, i);
}

int unpack_samples(const void* d_raw, int fmt, float2* d_out, int n, cudaStream_t st, LaunchCounter* lc)
{
    k_unpack<<<(n + 255) / 256, 256, 0, st>>>(d_raw, fmt, d_out, n);
    if (lc) lc->n++;
    CSDR_CK(cudaGetLastError());
    return CUTESDR_OK;
}

__global__ void k_scale_prefix(float2* x, const float* gain, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { float g = gain[i]; x[i].x *= g; x[i].y *= g; }
}

int apply_nco_startup_gain(float2* d_x, long long stream_pos, int n, cudaStream_t st, LaunchCounter* lc)
{
    if (stream_pos >= kNcoStartup || n <= 0) return CUTESDR_OK;
    static float h_gain[kNcoStartup];
    static std::once_flag once;
    std::call_once(once, []() {
        // a_0 = 1, a_{n+1} = a_n (1.95 - a_n^2): |Osc| seen by stream sample n
        double a = 1.0;
        const double steady = sqrt(0.95);
        for (int i = 0; i < kNcoStartup; i++) { h_gain[i] = (float)(a / steady); a = a * (1.95 - a * a); }
    });
    int m = (int)std::min<long long>(n, kNcoStartup - stream_pos);
    float* d_gain = nullptr;
    CSDR_CK(cudaMallocAsync(&d_gain, m * sizeof(float), st));
    CSDR_CK(cudaMemcpyAsync(d_gain, h_gain + stream_pos, m * sizeof(float), cudaMemcpyHostToDevice, st));
    k_scale_prefix<<<(m + 127) / 128, 128, 0, st>>>(d_x, d_gain, m);
    if (lc) lc->n++;
    CSDR_CK(cudaGetLastError());
    CSDR_CK(cudaFreeAsync(d_gain, st));
    return CUTESDR_OK;
}

// ------------------------------------------------------------------------------------------
// Decimator
// ------------------------------------------------------------------------------------------
int Decimator::read_timing(double* ms_total, long long* launches)
{
    CSDR_CK(cudaStreamSynchronize(st_));
    double gap = 0.0;
    for (size_t i = 0; i < ev_used_; i++) {
        float ms = 0.f;
        CSDR_CK(cudaEventElapsedTime(&ms, ev_pool_[i].first, ev_pool_[i].second));
        k1_ms_ += ms;
        k1_n_++;
        if (i + 1 < ev_used_) {
            CSDR_CK(cudaEventElapsedTime(&ms, ev_pool_[i].second, ev_pool_[i + 1].first));
            gap += ms;
        }
    }
    if (tun_.debug_timing && ev_used_ > 1)
        fprintf(stderr, "[cutesdr] kernel-1: %zu launches, mean %.3f ms, mean gap to next launch %.3f ms\n", ev_used_,
                k1_ms_ / (double)k1_n_, gap / (double)(ev_used_ - 1));
    ev_used_ = 0;
    if (ms_total) *ms_total = k1_ms_;
    if (launches) *launches = k1_n_;
    k1_ms_ = 0.0;
    k1_n_ = 0;
    return CUTESDR_OK;
}

int Decimator::set_overlap(bool on)
{
    if (on && !st_hb_) {
        int lo = 0, hi = 0;
        CSDR_CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CSDR_CK(cudaStreamCreateWithPriority(&st_hb_, cudaStreamNonBlocking, hi));
        for (auto& e : ev_k2_) CSDR_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    overlap_ = on && (int)lens_.size() > k1_stages() && !tun_.no_overlap;
    return CUTESDR_OK;
}

int Decimator::wait_before_output(cudaEvent_t ev)
{
    cudaStream_t s = (overlap_ && st_hb_) ? st_hb_ : st_;
    CSDR_CK(cudaStreamWaitEvent(s, ev, 0));
    return CUTESDR_OK;
}

int Decimator::join_main()
{
    if (overlap_ && blocks_run_ > 0) CSDR_CK(cudaStreamWaitEvent(st_, ev_done_, 0));
    return CUTESDR_OK;
}

Decimator::~Decimator()
{
    if (st_hb_) { cudaStreamSynchronize(st_hb_); cudaStreamDestroy(st_hb_); }
    if (ev_k1_) cudaEventDestroy(ev_k1_);
    if (ev_done_) cudaEventDestroy(ev_done_);
    for (auto& e : ev_k2_) if (e) cudaEventDestroy(e);
    for (auto& p : ev_pool_) { cudaEventDestroy(p.first); cudaEventDestroy(p.second); }
    cudaFree(d_nco_);
    cudaFree(d_tc_coef_);
    cudaFree(d_tc_coef16_);
    cudaFree(d_phase_[0]);
    cudaFree(d_phase_[1]);
    for (float2* p : d_stage_) cudaFree(p);
    cudaFree(d_ring_);
}

static int tap_offset_for(int len)
{
    for (int k = 1; k < CSDR_HB_NUM_KINDS; k++)
        if (csdr_hb_len[k] == len) return csdr_hb_tap_off[k];
    return 0;
}

int Decimator::init(int nch, double in_rate, double max_bw, int block_len, cudaStream_t st, LaunchCounter* lc)
{
    if (nch <= 0 || block_len <= 0) { set_error("Decimator::init: bad sizes"); return CUTESDR_E_ARG; }
    nch_ = nch;
    stride_ = nch >= 32 ? round_up(nch, 32) : next_pow2(nch);
    in_rate_ = in_rate;
    block_len_ = block_len;
    st_ = st;
    lc_ = lc;
    tun_ = Tuning::from_env();
    out_rate_ = plan_stages(in_rate, max_bw, lens_);
    if ((int)lens_.size() > kMaxStages) { set_error("too many decimation stages (%d)", (int)lens_.size()); return CUTESDR_E_ARG; }
    if (block_len % (1 << lens_.size()) != 0) {
        set_error("block length %d is not a multiple of 2^%d stages (dsp/downconvert.cpp:182-183)", block_len, (int)lens_.size());
        return CUTESDR_E_ARG;
    }
    ncic_ = 0;
    while (ncic_ < (int)lens_.size() && lens_[ncic_] == 3 && ncic_ < 6) ncic_++;
    // kernel 1T (tensor cores): the ladder starts with >= 4 CIC3 stages and every legal block length is a whole
    // number of 256-sample units (>= 8 stages), so the CUDA-core kernel never has to stand in for it
    tc_ = ncic_ >= 4 && ncic_ <= 6 && lens_.size() >= 8 && block_len % 256 == 0 && block_len >= 4096 && !tun_.no_tc;
    // up to kFuseHb 11-tap half-bands that follow the CICs run inside kernel 1 as well. Kernel 1T's epilogue has
    // the issue slots for more of them (kFuseHbTc), as long as the segment priming stays small.
    nhbf_ = 0;
    int fuse_max = tc_ ? kFuseHbTc : kFuseHb;
    if (tun_.fuse_hb >= 0) fuse_max = std::max(0, std::min(tc_ ? 3 : 2, tun_.fuse_hb));     // tuning aid
    while (nhbf_ < fuse_max && ncic_ + nhbf_ < (int)lens_.size() && lens_[ncic_ + nhbf_] == 11) nhbf_++;
    if (tc_) {
        while (nhbf_ > 0 && (16 * tc_pre(ncic_ - 4, nhbf_) + 32 > kHaloMax || tc_pre(ncic_ - 4, nhbf_) > 96)) nhbf_--;
        if (16 * tc_pre(ncic_ - 4, nhbf_) + 32 > kHaloMax) tc_ = false;
    }
    n_out_ = block_len >> lens_.size();
    if (n_out_ > kDecRing - kFirFft) { set_error("decimated block of %d samples exceeds the FIR ring", n_out_); return CUTESDR_E_ARG; }
    CSDR_TRY(upload_taps());

    h_nco_.assign(stride_, NcoDev{0ull, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f});
    CSDR_CK(cudaMalloc(&d_nco_, stride_ * sizeof(NcoDev)));
    for (int k = 0; k < 2; k++) {
        CSDR_CK(cudaMalloc(&d_phase_[k], stride_ * sizeof(unsigned long long)));
        CSDR_CK(cudaMemsetAsync(d_phase_[k], 0, stride_ * sizeof(unsigned long long), st_));
    }
    const int nhb = (int)lens_.size() - k1_stages();
    for (int s = 0; s < nhb; s++) {
        int n_rows = block_len >> (k1_stages() + s);       // rows this ring receives per block
        // history a fused pass over stages s.. reaches back before the block (k_hb_tail recomputes the later stages'
        // history from this ring instead of keeping their rings)
        long long hist = 0;
        for (int j = nhb - 1; j >= s; j--) hist = std::min<long long>(2 * hist + (lens_[k1_stages() + j] - 1), 8192);
        // ring 0 keeps two blocks so kernel 1 of block k+1 can fill it while kernel 2 still reads block k
        int rows = next_pow2((long long)n_rows * (s == 0 ? 2 : 1) + 64 + ((nhb - s <= 4 || s == 0) ? hist + kH3Box : 0));
        float2* p = nullptr;
        size_t bytes = (size_t)rows * stride_ * sizeof(float2);
        CSDR_CK(cudaMalloc(&p, bytes));
        CSDR_CK(cudaMemsetAsync(p, 0, bytes, st_));
        d_stage_.push_back(p);
        stage_rows_.push_back(rows);
        stage_base_.push_back(0);
    }
    // kernel 2s (register-streaming pass over the first three stages) applies when they are 11, 11 and 11 or 15 taps long,
    // which is every ladder of SURVEY's configs once kernel 1 has taken its share
    hs_stages_ = 0;
    if (nhb >= 3 && stride_ % kH3Ch == 0 && !tun_.no_hbstream && (block_len >> k1_stages()) % kH3Box == 0 &&
        lens_[k1_stages()] == 11 && lens_[k1_stages() + 1] == 11 && (lens_[k1_stages() + 2] == 11 || lens_[k1_stages() + 2] == 15))
        hs_stages_ = 3;
    if (hs_stages_ > 0) {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CSDR_CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { set_error("cuTensorMapEncodeTiled is not available in this driver"); return CUTESDR_E_CUDA; }
        // stage ring 0 as a 2-D tensor of floats: x = 2 floats per channel (row = 2 stride floats), y = ring rows
        const cuuint64_t dims[2] = {(cuuint64_t)2 * stride_, (cuuint64_t)stage_rows_[0]};
        const cuuint64_t strides[1] = {(cuuint64_t)stride_ * sizeof(float2)};
        const cuuint32_t box[2] = {2 * kH3Ch, kH3Box};
        const cuuint32_t estr[2] = {1, 1};
        static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
        const CUresult r = reinterpret_cast<EncodeFn>(fn)(reinterpret_cast<CUtensorMap*>(tmap0_), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_stage_[0], dims,
                                                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CUTESDR_E_CUDA; }
        CSDR_CK(cudaFuncSetAttribute(k_hb3r<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, kH3Smem));
        CSDR_CK(cudaFuncSetAttribute(k_hb3r<15>, cudaFuncAttributeMaxDynamicSharedMemorySize, kH3Smem));
    }
    size_t rbytes = (size_t)stride_ * kDecRing * sizeof(float2);
    CSDR_CK(cudaMalloc(&d_ring_, rbytes));
    CSDR_CK(cudaMemsetAsync(d_ring_, 0, rbytes, st_));
    CSDR_CK(cudaEventCreateWithFlags(&ev_k1_, cudaEventDisableTiming));
    CSDR_CK(cudaEventCreateWithFlags(&ev_done_, cudaEventDisableTiming));

    if (!tc_) {
        // Time tile. Every CTA costs ~(tile + halo) samples per lane, and the grid runs in waves of
        // (SMs x resident CTAs): pick the tile count that minimises waves x (tile + halo), i.e. avoid a
        // mostly-empty last wave and keep the halo small, within the shared memory that still allows the
        // same residency.
        const int B = k1_body(ncic_);
        const int Q = std::max((1 << ncic_) << nhbf_, B);
        const int H = k1_halo(ncic_, nhbf_);
        if (H > kHaloMax) { set_error("kernel-1 halo %d exceeds kHaloMax", H); return CUTESDR_E_ARG; }
        const int cta_threads = std::min(256, round_up(stride_, 32));
        const int chan_blocks = (stride_ + cta_threads - 1) / cta_threads;
        int dev = 0, sms = 148;
        CSDR_CK(cudaGetDevice(&dev));
        CSDR_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        K1Fn fn = k1_kernel(ncic_, nhbf_);
        CSDR_CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        double best_cost = 1e300;
        int best_tl = Q;
        std::map<int, int> occ_cache;       // smem KB -> resident CTAs per SM
        const int max_tiles = std::max(1, block_len / std::max(Q, 4 * H));
        for (int tiles = 1; tiles <= max_tiles; tiles++) {
            int tl = (block_len + tiles - 1) / tiles;
            tl = (tl + Q - 1) / Q * Q;
            const size_t smem = (size_t)(tl + H) * sizeof(float2);
            if (smem > 200 * 1024) continue;
            const int real_tiles = (block_len + tl - 1) / tl;
            const int key = (int)(smem >> 10);
            int occ;
            auto it = occ_cache.find(key);
            if (it != occ_cache.end()) occ = it->second;
            else {
                CSDR_CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, cta_threads, smem));
                occ_cache[key] = occ;
            }
            if (occ < 1) continue;
            const long long grid = (long long)real_tiles * chan_blocks;
            const long long slots = (long long)sms * occ;
            const long long waves = (grid + slots - 1) / slots;
            // fewer resident warps hide latency worse: charge a residency penalty below 16 warps/SM
            const double warps = (double)std::min<long long>(grid, slots) / sms * (cta_threads / 32.0);
            const double penalty = warps >= 16.0 ? 1.0 : (16.0 / std::max(warps, 1.0));
            const double cost = (double)waves * (tl + H) * (warps >= 16.0 ? 1.0 : std::min(penalty, 4.0) * 0.5 + 0.5);
            if (cost < best_cost * 0.999) { best_cost = cost; best_tl = tl; }
        }
        tile_len_ = best_tl;
        if (tun_.tile > 0) tile_len_ = std::max(Q, tun_.tile / Q * Q);                    // tuning aid
        if (tun_.debug_timing) {
            int occ = 0;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, cta_threads, (size_t)(tile_len_ + H) * sizeof(float2));
            fprintf(stderr, "[cutesdr] kernel-1 <%d,%d>: tile %d + halo %d, grid %d x %d, %d CTAs/SM\n", ncic_, nhbf_, tile_len_, H,
                    (block_len + tile_len_ - 1) / tile_len_, chan_blocks, occ);
        }
        // kernel 1T: the ladder starts with >= 4 CIC3 stages and the block is a whole number of 256-sample units
    }
    if (tc_) {
        int dev = 0, sms = 148;
        CSDR_CK(cudaGetDevice(&dev));
        CSDR_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        tc_groups_ = (stride_ + 127) / 128;
        CSDR_CK(cudaMalloc(&d_tc_coef_, (size_t)tc_groups_ * 128 * 192 * sizeof(float)));
        // two forms of the A table: tf32 hi/lo for float32 / int24 blocks, fp16 hi/lo for int16 blocks (see k_mix_tc)
        CSDR_CK(cudaMalloc(&d_tc_coef16_, (size_t)tc_groups_ * 128 * 192 * sizeof(float)));
        CSDR_CK(cudaFuncSetAttribute(k1t_kernel(ncic_ - 4, nhbf_, false), cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
        CSDR_CK(cudaFuncSetAttribute(k1t_kernel(ncic_ - 4, nhbf_, true), cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
        // one persistent CTA per SM, kTcSegs interleaved time segments per CTA; every segment pays PRE priming
        // outputs and rounds up to whole MMA tiles
        const int pre = tc_pre(ncic_ - 4, nhbf_);
        {
            // tc_spare SMs are left to the burst chain's packed sequential kernels (see post.cu)
            const int ctas_x = std::max(1, (sms - tun_.tc_spare) / tc_groups_);
            int sl = (block_len + kTcSegs * ctas_x - 1) / (kTcSegs * ctas_x);
            tc_seg_len_ = std::max(256, (sl + 255) / 256 * 256);
        }
        if (tun_.tc_seg > 0) tc_seg_len_ = std::max(256, tun_.tc_seg / 256 * 256);           // tuning aid
        tc_dirty_ = true;
        if (tun_.debug_timing)
            fprintf(stderr, "[cutesdr] kernel-1T <%d,%d>: segment %d (+%d priming), grid %d x %d\n", ncic_ - 4, nhbf_, tc_seg_len_,
                    16 * pre, ((block_len + tc_seg_len_ - 1) / tc_seg_len_ + kTcSegs - 1) / kTcSegs, tc_groups_);
    }
    // Opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: set it for the device this
    // object lives on, every time (a process may hold banks on several GPUs).
    CSDR_CK(cudaFuncSetAttribute(k_hb11_chain<2, kHbc2T, kHbcCh>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)HbcCfg<2, kHbc2T>::smem_rows() * kHbcCh * sizeof(float2))));
    CSDR_CK(cudaFuncSetAttribute(k_hb11_chain<3, kHbc3T, kHbcCh>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)((size_t)HbcCfg<3, kHbc3T>::smem_rows() * kHbcCh * sizeof(float2))));
    CSDR_CK(cudaFuncSetAttribute(k_hb_tail<kHbcCh>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    dirty_ = true;
    return CUTESDR_OK;
}

void Decimator::set_frequency(int i, double nco_freq)
{
    // CDownConvert::SetFrequency, dsp/downconvert.cpp:98-107. The phase increment is kept as a
    // 64-bit binary fraction of a turn so the tile-start phase of every time tile is exact.
    long double turns = (long double)nco_freq / (long double)in_rate_;
    turns -= floorl(turns);
    unsigned long long inc = (unsigned long long)(turns * 18446744073709551616.0L);
    const int B = k1_body(ncic_);
    const double a1 = kTwoPi * (double)((long double)inc / 18446744073709551616.0L);
    const unsigned long long incB = inc * (unsigned long long)B;
    const double aB = kTwoPi * (double)((long double)incB / 18446744073709551616.0L);
    NcoDev& n = h_nco_[i];
    n.inc = inc;
    n.w1c = (float)cos(a1); n.w1s = (float)sin(a1);
    n.wgc = (float)cos(aB); n.wgs = (float)sin(aB);
    const double a16 = kTwoPi * (double)((long double)(inc * 16ull) / 18446744073709551616.0L);
    n.wtc = (float)cos(a16); n.wts = (float)sin(a16);
    dirty_ = true;
    tc_dirty_ = true;
}

int Decimator::reset_channel(int i)
{
    if (i < 0 || i >= stride_) { set_error("Decimator::reset_channel: slot %d of %d", i, stride_); return CUTESDR_E_ARG; }
    if (overlap_) CSDR_TRY(join_main());
    for (size_t s = 0; s < d_stage_.size(); s++)      // time-major rings: one float2 column
        CSDR_CK(cudaMemset2DAsync(d_stage_[s] + i, (size_t)stride_ * sizeof(float2), 0, sizeof(float2), stage_rows_[s], st_));
    CSDR_CK(cudaMemsetAsync(d_ring_ + (size_t)i * kDecRing, 0, (size_t)kDecRing * sizeof(float2), st_));
    for (int k = 0; k < 2; k++) CSDR_CK(cudaMemsetAsync(d_phase_[k] + i, 0, sizeof(unsigned long long), st_));
    h_nco_[i] = NcoDev{0ull, 1.f, 0.f, 1.f, 0.f, 1.f, 0.f};
    dirty_ = true;
    tc_dirty_ = true;
    return CUTESDR_OK;
}

int Decimator::upload_dirty()
{
    if (!dirty_) return CUTESDR_OK;
    // pinned snapshot: no stream synchronisation, and h_nco_ may be modified again as soon as this returns
    CSDR_TRY(stage_.upload(d_nco_, h_nco_.data(), stride_ * sizeof(NcoDev), st_));
    dirty_ = false;
    return CUTESDR_OK;
}

int Decimator::run_block(const void* d_x, const float2* halo_cur, float2* halo_next, int L, int fmt)
{
    CSDR_TRY(upload_dirty());
    if (L < 0) L = block_len_;
    if (L == 0) return CUTESDR_OK;
    if (L > block_len_ || L % (1 << lens_.size()) != 0) {
        set_error("Decimator::run_block: length %d (capacity %d, must be a multiple of %d)", L, block_len_, 1 << lens_.size());
        return CUTESDR_E_ARG;
    }
    const int nhb = (int)lens_.size() - k1_stages();
    OutDesc od;
    if (nhb > 0) {
        od.p = d_stage_[0];
        od.mask = (unsigned)(stage_rows_[0] - 1);
        od.stride = stride_;
        od.transposed = 0;
        od.base = stage_base_[0];
    } else {
        od.p = d_ring_;
        od.mask = 0;
        od.stride = stride_;
        od.transposed = 1;
        od.base = total_out_;
    }
    // steady-state oscillator amplitude sqrt(0.95) and the folded .125 per CIC3 stage
    const float scale = (float)(sqrt(0.95) * ldexp(1.0, -3 * ncic_));
    const unsigned long long* pc = d_phase_[phase_cur_];
    unsigned long long* pn = d_phase_[phase_cur_ ^ 1];
    if (overlap_ && blocks_run_ >= 2) CSDR_CK(cudaStreamWaitEvent(st_, ev_k2_[(blocks_run_ - 2) & 3], 0));
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;
    if (timing_ && ev_used_ < 8192) {
        if (ev_used_ == ev_pool_.size()) {
            cudaEvent_t a, b2;
            CSDR_CK(cudaEventCreate(&a));
            CSDR_CK(cudaEventCreate(&b2));
            ev_pool_.push_back({a, b2});
        }
        ev_a = ev_pool_[ev_used_].first;
        ev_b = ev_pool_[ev_used_].second;
        ev_used_++;
        CSDR_CK(cudaEventRecord(ev_a, st_));
    }
    const int Q = std::max((1 << ncic_) << nhbf_, k1_body(ncic_));
    if (tc_ && L % 256 == 0) {
        if (tc_dirty_) {
            const int n = tc_groups_ * 128 * 48;
            k_tc_coeffs<<<(n + 255) / 256, 256, 0, st_>>>(d_nco_, stride_, tc_groups_, d_tc_coef_);
            k_tc_coeffs_f16<<<(n / 2 + 255) / 256, 256, 0, st_>>>(d_nco_, stride_, tc_groups_, reinterpret_cast<uint32_t*>(d_tc_coef16_));
            lc_->n += 2;
            CSDR_CK(cudaGetLastError());
            tc_dirty_ = false;
        }
        tc_f16_ = fmt == 1;            // int16 blocks take the exact fp16 form
        const int sl = std::min(tc_seg_len_, L);
        dim3 grid(((L + sl - 1) / sl + kTcSegs - 1) / kTcSegs, tc_groups_);
        k1t_kernel(ncic_ - 4, nhbf_, tc_f16_)<<<grid, kTcThreads, kTcSmem, st_>>>(d_x, fmt, halo_cur, halo_next, L, sl, tc_f16_ ? d_tc_coef16_ : d_tc_coef_,
                                                                          d_nco_, pc, pn, stride_, od, scale);
    } else if (L % Q == 0) {
        const int threads = std::min(256, round_up(stride_, 32));
        dim3 grid((L + tile_len_ - 1) / tile_len_, (stride_ + threads - 1) / threads);
        const int H = k1_halo(ncic_, nhbf_);
        size_t smem = (size_t)(tile_len_ + H) * sizeof(float2);
        k1_kernel(ncic_, nhbf_)<<<grid, threads, smem, st_>>>(d_x, fmt, halo_cur, halo_next, L, tile_len_, d_nco_, pc, pn, stride_, od,
                                                               scale);
    } else {
        k_mix_cic_generic<<<(stride_ + 63) / 64, 64, 0, st_>>>(d_x, fmt, halo_cur, halo_next, L, ncic_, nhbf_, d_nco_, pc, pn, stride_, od, scale);
    }
    lc_->n++;
    CSDR_CK(cudaGetLastError());
    if (ev_b) CSDR_CK(cudaEventRecord(ev_b, st_));
    phase_cur_ ^= 1;
    cudaStream_t s2 = st_;
    if (overlap_) {
        CSDR_CK(cudaEventRecord(ev_k1_, st_));
        CSDR_CK(cudaStreamWaitEvent(st_hb_, ev_k1_, 0));
        s2 = st_hb_;
    }

    int s_first = 0;
    if (hs_stages_ == 3 && (L >> k1_stages()) % kH3Box == 0 && stage_base_[0] % kH3Box == 0) {
        // kernel 2s: stages 0-2 in one register-streaming pass
        const int L2 = lens_[k1_stages() + 2];
        const int n2 = L >> (k1_stages() + 3);
        int dev = 0, sms = 148;
        CSDR_CK(cudaGetDevice(&dev));
        CSDR_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        const int groups = stride_ / kH3Ch;
        // ~hs_ctas x 2 warps per SM in total; a slice is at least 16 outputs (128 rows) so the ~100 priming rows stay a fraction
        // -- rounded DOWN to whole CTAs: every CTA lives as long as the kernel, so the grid must fit in ONE wave
        // (sms * hs_ctas resident CTAs; 448 CTAs on 444 slots ran as two waves and doubled the pass)
        int slices = std::max(1, std::min(n2 / 16, (sms * tun_.hs_ctas * kH3Warps) / groups));
        slices = std::max(kH3Warps, slices / kH3Warps * kH3Warps);
        const int per_slice = (n2 + slices - 1) / slices;
        OutDesc o2;
        if (3 < nhb) {
            o2.p = d_stage_[3];
            o2.mask = (unsigned)(stage_rows_[3] - 1);
            o2.stride = stride_;
            o2.transposed = 0;
            o2.base = stage_base_[3];
        } else {
            o2.p = d_ring_;
            o2.mask = 0;
            o2.stride = stride_;
            o2.transposed = 1;
            o2.base = total_out_;
        }
        dim3 grid(((n2 + per_slice - 1) / per_slice + kH3Warps - 1) / kH3Warps, groups);
        const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(tmap0_);
        if (L2 == 11) k_hb3r<11><<<grid, 32 * kH3Warps, kH3Smem, s2>>>(tm, (unsigned)(stage_rows_[0] - 1), stage_base_[0], n2, per_slice, tap_offset_for(11), o2);
        else k_hb3r<15><<<grid, 32 * kH3Warps, kH3Smem, s2>>>(tm, (unsigned)(stage_rows_[0] - 1), stage_base_[0], n2, per_slice, tap_offset_for(15), o2);
        lc_->n++;
        CSDR_CK(cudaGetLastError());
        for (int s = 0; s < 3; s++) stage_base_[s] += (L >> (k1_stages() + s));
        s_first = 3;
    }
    if (s_first == 0) {
        // leading run of 11-tap stages -> one fused pass
        int nchain = 0;
        while (nchain < 3 && nchain < nhb && lens_[k1_stages() + nchain] == 11) nchain++;
        if (nchain >= 2 && stride_ % 32 == 0 && !tun_.no_hbchain) {
            const int n_in = L >> k1_stages();
            const int n_out = n_in >> nchain;
            OutDesc o2;
            if (nchain < nhb) {
                o2.p = d_stage_[nchain];
                o2.mask = (unsigned)(stage_rows_[nchain] - 1);
                o2.stride = stride_;
                o2.transposed = 0;
                o2.base = stage_base_[nchain];
            } else {
                o2.p = d_ring_;
                o2.mask = 0;
                o2.stride = stride_;
                o2.transposed = 1;
                o2.base = total_out_;
            }
            const unsigned mask0 = (unsigned)(stage_rows_[0] - 1);
            constexpr int CH = kHbcCh;
            const int chan_blocks = stride_ / CH;
            if (nchain == 2) {
                constexpr int T = kHbc2T;
                const size_t smem = (size_t)HbcCfg<2, T>::smem_rows() * CH * sizeof(float2);
                dim3 grid((n_out + T - 1) / T, chan_blocks);
                k_hb11_chain<2, T, CH><<<grid, 256, smem, s2>>>(d_stage_[0], mask0, stage_base_[0], stride_, n_out, o2);
            } else {
                constexpr int T = kHbc3T;
                const size_t smem = (size_t)HbcCfg<3, T>::smem_rows() * CH * sizeof(float2);
                dim3 grid((n_out + T - 1) / T, chan_blocks);
                k_hb11_chain<3, T, CH><<<grid, 256, smem, s2>>>(d_stage_[0], mask0, stage_base_[0], stride_, n_out, o2);
            }
            lc_->n++;
            CSDR_CK(cudaGetLastError());
            for (int s = 0; s < nchain; s++) stage_base_[s] += (L >> (k1_stages() + s));
            s_first = nchain;
        }
    }
    {
        // the remaining stages in one launch when there are 2..4 of them and the channel rows are whole 128-byte lines
        const int ns = nhb - s_first;
        // opt-in (CUTESDR_HBTAIL=1): measured neutral inside the step (0.322 vs 0.320 ms, cfg4) -- the three small launches it
        // replaces already overlap with the burst chain, and its deep halo recomputes ~1.8x of the first two stages
        bool ok = ns >= 2 && ns <= 4 && stride_ % 16 == 0 && (tun_.hbtail || (s_first == 3 && hs_stages_ == 3 && !tun_.no_hbtail));
        for (int s = s_first; ok && s < nhb; s++) ok = lens_[k1_stages() + s] != 3;
        if (ok) {
            constexpr int CH = kHbcCh;
            TailCfg cfg;
            cfg.ns = ns;
            const int n_out = L >> lens_.size();
            // tile length: every CTA lives about as long as its c[0] input rows; pick the T that minimises waves x rows
            int dev = 0, sms = 148;
            CSDR_CK(cudaGetDevice(&dev));
            CSDR_CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            int T = 8;
            long long best = -1;
            for (int cand = 64; cand >= 8; cand -= 8) {
                int c[5];
                c[ns] = cand;
                for (int j = ns - 1; j >= 0; j--) c[j] = 2 * c[j + 1] + lens_[k1_stages() + s_first + j] - 2;
                bool fits = true;
                for (int j = 0; j < ns; j++) fits &= (c[j + 1] * CH + 255) / 256 <= kTailMaxPerThread;
                const size_t bytes = (size_t)c[0] * CH * sizeof(float2);
                if (!fits || bytes > 160 * 1024) continue;
                const int per_sm = (int)std::min<size_t>(4, (227 * 1024) / (bytes + 1024));       // __launch_bounds__(256, 4)
                const long long ctas = (long long)((n_out + cand - 1) / cand) * (stride_ / CH);
                const long long waves = (ctas + (long long)sms * per_sm - 1) / ((long long)sms * per_sm);
                const long long cost = waves * c[0];
                if (best < 0 || cost < best) { best = cost; T = cand; }
            }
            cfg.c[ns] = T;
            for (int j = ns - 1; j >= 0; j--) {
                cfg.len[j] = lens_[k1_stages() + s_first + j];
                cfg.toff[j] = tap_offset_for(cfg.len[j]);
                cfg.c[j] = 2 * cfg.c[j + 1] + cfg.len[j] - 2;
            }
            const size_t smem = (size_t)cfg.c[0] * CH * sizeof(float2);
            // rows of history the first tile reaches back into the input ring (it persists between blocks)
            long long hist = 0;
            for (int j = ns - 1; j >= 0; j--) hist = 2 * hist + (cfg.len[j] - 1);
            const int n_in0 = L >> (k1_stages() + s_first);
            if (best >= 0 && smem <= 160 * 1024 && hist + n_in0 <= stage_rows_[s_first]) {
                OutDesc o2;
                o2.p = d_ring_;
                o2.mask = 0;
                o2.stride = stride_;
                o2.transposed = 1;
                o2.base = total_out_;
                dim3 grid((n_out + T - 1) / T, stride_ / CH);
                k_hb_tail<CH><<<grid, 256, smem, s2>>>(d_stage_[s_first], (unsigned)(stage_rows_[s_first] - 1), stage_base_[s_first], stride_,
                                                       n_out, cfg, o2);
                lc_->n++;
                CSDR_CK(cudaGetLastError());
                for (int s = s_first; s < nhb; s++) stage_base_[s] += (L >> (k1_stages() + s));
                s_first = nhb;
            }
        }
    }
    for (int s = s_first; s < nhb; s++) {
        const int N = lens_[k1_stages() + s];
        const int n_in = L >> (k1_stages() + s);
        const int n_out = n_in >> 1;
        OutDesc o2;
        if (s + 1 < nhb) {
            o2.p = d_stage_[s + 1];
            o2.mask = (unsigned)(stage_rows_[s + 1] - 1);
            o2.stride = stride_;
            o2.transposed = 0;
            o2.base = stage_base_[s + 1];
        } else {
            o2.p = d_ring_;
            o2.mask = 0;
            o2.stride = stride_;
            o2.transposed = 1;
            o2.base = total_out_;
        }
        long long work = (long long)n_out * stride_;
        int blocks = (int)((work + 127) / 128);      // small CTAs slot in beside kernel 1's resident CTAs
        k_halfband<<<blocks, 128, 0, s2>>>(d_stage_[s], (unsigned)(stage_rows_[s] - 1), stage_base_[s], stride_, n_out, N,
                                           tap_offset_for(N), o2);
        lc_->n++;
        CSDR_CK(cudaGetLastError());
        stage_base_[s] += n_in;
    }
    if (overlap_) CSDR_CK(cudaEventRecord(ev_k2_[blocks_run_ & 3], s2));
    CSDR_CK(cudaEventRecord(ev_done_, s2));
    blocks_run_++;
    total_out_ += L >> lens_.size();
    return CUTESDR_OK;
}

}  // namespace csdr
