// post.cu -- kernel group 4: S-meter, AGC and the demodulators on 1024-sample bursts.
//
// Everything that is pointwise or feed-forward in time runs time-parallel, one CTA per channel:
//   k_post_pre  : S-meter dB values, AGC log-magnitudes and their exact sliding-window maximum
//                 (doubling in shared memory; the reference's peak tracker IS a windowed max)
//   k_post_mid  : AGC gain law (pow) x delayed signal, then envelope / phase angle per mode
//   k_post_fir  : Kaiser FIRs (AM post filter, FM squelch high-pass), squelch decision, biquad
// Only the true recurrences stay sequential, one LANE per channel:
//   k_post_seq1 : S-meter attack/decay poles, AGC attack/decay averagers + hang timer
//   k_post_seq2 : AM/SAM DC blockers and the SAM/FM second-order PLLs. The PLL detector
//                 atan2(Im(x e^{-j phi}), Re(..)) equals wrap(arg(x) - phi); arg(x) is computed
//                 time-parallel in k_post_mid, so the sequential loop is adds, a wrap and a clamp.
// Recurrent state is double precision (averagers with 1e4-sample time constants would sit near
// -85 dB in float32); burst data is float32. All rows are channel-major [c][row].
#include "post.cuh"
#include <type_traits>

namespace csdr {

// ------------------------------------------------------------------------------------------
// host: Kaiser FIR designs (dsp/fir.cpp:173-367, 414-432)
// ------------------------------------------------------------------------------------------
static double bessel_i0(double x)
{
    double x2 = x / 2.0, sum = 1.0, ds = 1.0, di = 1.0, t;
    do {
        t = x2 / di;
        t *= t;
        ds *= t;
        sum += ds;
        di += 1.0;
    } while (ds >= 1e-9 * sum);
    return sum;
}

static double kaiser_beta(double astop)
{
    if (astop < 20.96) return 0;
    if (astop >= 50.0) return .1102 * (astop - 8.71);
    return .5842 * pow((astop - 20.96), 0.4) + .07886 * (astop - 20.96);
}

int design_kaiser_lp(double scale, double astop, double fpass, double fstop, double fs, double* coef)
{
    const double nfp = fpass / fs, nfs = fstop / fs, nfc = (nfs + nfp) / 2.0;
    const double beta = kaiser_beta(astop);
    int ntaps = (astop - 8.0) / (2.285 * kTwoPi * (nfs - nfp)) + 1;
    if (ntaps > kFirMax) ntaps = kFirMax;
    if (ntaps < 3) ntaps = 3;
    const double centre = .5 * (double)(ntaps - 1);
    const double izb = bessel_i0(beta);
    for (int n = 0; n < ntaps; n++) {
        double x = (double)n - centre, c;
        if ((double)n == centre) c = 2.0 * nfc;
        else c = sin(kTwoPi * x * nfc) / (kPi * x);
        x = ((double)n - ((double)ntaps - 1.0) / 2.0) / (((double)ntaps - 1.0) / 2.0);
        coef[n] = scale * c * bessel_i0(beta * sqrt(1 - (x * x))) / izb;
    }
    return ntaps;
}

int design_kaiser_hp(double scale, double astop, double fpass, double fstop, double fs, double* coef)
{
    const double nfp = fpass / fs, nfs = fstop / fs, nfc = (nfs + nfp) / 2.0;
    const double beta = kaiser_beta(astop);
    int ntaps = (astop - 8.0) / (2.285 * kTwoPi * (nfp - nfs)) + 1;
    if (ntaps > (kFirMax - 1)) ntaps = kFirMax - 1;
    if (ntaps < 3) ntaps = 3;
    ntaps |= 1;
    const double izb = bessel_i0(beta);
    const double centre = .5 * (double)(ntaps - 1);
    for (int n = 0; n < ntaps; n++) {
        double x = (double)n - (double)(ntaps - 1) / 2.0, c;
        if ((double)n == centre) c = 1.0 - 2.0 * nfc;
        else c = (sin(kPi * x) / (kPi * x) - sin(kTwoPi * x * nfc) / (kPi * x));
        x = ((double)n - ((double)ntaps - 1.0) / 2.0) / (((double)ntaps - 1.0) / 2.0);
        coef[n] = scale * c * bessel_i0(beta * sqrt(1 - (x * x))) / izb;
    }
    return ntaps;
}

// ------------------------------------------------------------------------------------------
// device state layout: scalars are struct-of-arrays [field][stride]
// ------------------------------------------------------------------------------------------
enum { P_AGC_ON, P_AGC_HANG, P_KNEE, P_GAIN_SLOPE, P_FIXED_GAIN, P_MANUAL_GAIN, P_A_RISE, P_A_FALL, P_D_RISE,
       P_D_FALL, P_HANG_TIME, P_SQ_THRESH, P_NTAPS, P_COUNT };
enum { S_SM_ATT, S_SM_DEC, S_SM_AVE, S_SM_PEAK, S_AGC_ATT, S_AGC_DEC, S_Z1, S_Y1, S_PHASE, S_FREQ, S_FM_DC,
       S_SQ_AVE, S_LP_W1, S_LP_W2, S_LP_W1N, S_LP_W2N, S_COUNT };
enum { I_AGC_HANGT, I_SQUELCHED, I_COUNT };
enum { R_AGC = 1, R_DEMOD = 2, R_FIR = 4, R_SMETER = 8 };

constexpr int kHist = kFirMax - 1;

struct PostBufs {
    float2* y; int y_row;          // [c][kYHist + max_n]
    double* magh;                  // [c][kAgcBuf]
    double* smag; double* peak; float2* z; double* u; double* th; int row;   // [c][max_n]
    double* v; double* v2; int v_row;   // [c][kHist + max_n]
    const double* sam_taps;        // [2][kFirMax]
    const double* par; const double* taps; const int* mode; int* reset;
    const double* qpow;            // [max_n + 1] powers of (1 - squelch alpha)
    double* state; int* istate;
    int nch, stride;
};

#define PAR(f) b.par[(size_t)(f) * b.stride + c]
#define ST(f) b.state[(size_t)(f) * b.stride + c]
#define IST(f) b.istate[(size_t)(f) * b.stride + c]

// ---- state (re)initialisation for channels whose reset flags are raised; CTA per channel
__global__ void __launch_bounds__(128) k_post_reset(PostBufs b)
{
    const int c = blockIdx.x;
    const int rf = b.reset[c];
    if (!rf) return;
    if (rf & R_AGC) {       // dsp/agc.cpp:121-136: delay line zero, magnitude window -16, averagers -5
        for (int i = threadIdx.x; i < kYHist; i += blockDim.x) b.y[(size_t)c * b.y_row + i] = make_float2(0.f, 0.f);
        for (int i = threadIdx.x; i < kAgcBuf; i += blockDim.x) b.magh[(size_t)c * kAgcBuf + i] = -16.0;
    }
    if (rf & R_FIR) for (int i = threadIdx.x; i < kHist; i += blockDim.x) { b.v[(size_t)c * b.v_row + i] = 0.0; b.v2[(size_t)c * b.v_row + i] = 0.0; }
    if (threadIdx.x == 0) {
        if (rf & R_SMETER) { ST(S_SM_ATT) = -120.0; ST(S_SM_DEC) = -120.0; ST(S_SM_AVE) = 0.0; ST(S_SM_PEAK) = 0.0; }
        if (rf & R_AGC) { IST(I_AGC_HANGT) = 0; ST(S_AGC_DEC) = -5.0; ST(S_AGC_ATT) = -5.0; }
        if (rf & R_DEMOD) {
            ST(S_Z1) = 0.0; ST(S_Y1) = 0.0; ST(S_PHASE) = 0.0; ST(S_FREQ) = 0.0; ST(S_FM_DC) = 0.0; ST(S_SQ_AVE) = 0.0;
            ST(S_LP_W1) = 0.0; ST(S_LP_W2) = 0.0;
            IST(I_SQUELCHED) = 1;
        }
        b.reset[c] = 0;
    }
}

// ---- S-meter dB, AGC log-magnitude and its sliding-window maximum; CTA per channel
// dynamic smem: 2 x (window - 1 + n) doubles -- only indices [kAgcBuf - (window-1), kAgcBuf + n) of the two work
// arrays are ever touched, so they are allocated compactly (3 -> 9 resident CTAs per SM at n = 1024)
__global__ void __launch_bounds__(256) k_post_pre(PostBufs b, int n, int window)
{
    extern __shared__ double sm_d[];
    const int c = blockIdx.x;
    const int hn = window - 1;
    const int lo = kAgcBuf - hn;
    const int len_cmp = hn + n;
    double* A = sm_d - lo;                 // A[lo] is sm_d[0]
    double* B = sm_d + len_cmp - lo;
    const int mode = b.mode[c];
    const float2* yrow = b.y + (size_t)c * b.y_row + kYHist;
    for (int i = threadIdx.x; i < hn; i += blockDim.x) A[lo + i] = b.magh[(size_t)c * kAgcBuf + i];
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const float2 xf = yrow[t];
        const double xr = xf.x, xi = xf.y;
        if (mode != POST_AGC_ONLY) {   // dsp/smeter.cpp:76: 10 log10(|x|^2 / 32767^2 + 1e-50)
            // The meter is displayed in 0.1 dB steps and checked to 0.02 dB: log2 of the double's mantissa in float32
            // (error < 1e-6 dB) plus its exponent, instead of a double-precision log10 per sample.
            const double p = (xr * xr + xi * xi) / (32767.0 * 32767.0) + 1e-50;
            const int hiw = __double2hiint(p);
            const int ex = ((hiw >> 20) & 0x7ff) - 1023;
            const float mant = (float)__hiloint2double((hiw & 0x000fffff) | 0x3ff00000, __double2loint(p));      // [1, 2)
            b.smag[(size_t)c * b.row + t] = 3.0102999566398120 * ((double)ex + (double)log2f(mant));
        }
        double mag = fabs(xr);
        const double mim = fabs(xi);
        if (mim > mag) mag = mim;
        A[kAgcBuf + t] = log10(mag + 3.2767e-4) - log10(32767.0);     // dsp/agc.cpp:197-201
    }
    __syncthreads();
    // carry the last window-1 magnitudes for the next burst (everyone has read the old ones)
    for (int i = threadIdx.x; i < hn; i += blockDim.x) b.magh[(size_t)c * kAgcBuf + i] = A[kAgcBuf + n - hn + i];
    // doubling: after k passes src[i] = max over [max(lo, i-2^k+1), i]
    double* src = A;
    double* dst = B;
    int len = 1;
    const int hi = kAgcBuf + n;
    while (len * 2 <= window) {
        for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
            const double a0 = src[i];
            const int j = i - len;
            dst[i] = (j >= lo) ? fmax(a0, src[j]) : a0;
        }
        __syncthreads();
        double* tmp = src; src = dst; dst = tmp;
        len <<= 1;
    }
    // window of `window` samples ending at i = two overlapping power-of-two windows
    // (dsp/agc.cpp:210-231 keeps exactly this maximum: new sample vs current peak, rescan when the
    // sample leaving the window was the peak)
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const int i = kAgcBuf + t;
        b.peak[(size_t)c * b.row + t] = fmax(src[i], src[i - window + len]);
    }
}

// ------------------------------------------------------------------------------------------
// Row streaming for the sequential kernels (lane = channel, every lane walks its own row).
// A plain per-lane load touches 32 different lines per warp instruction; at 8 LDG.128 + 8 STG.128 per 16 samples the
// warps sat in the LSU queue (ncu: stall_lg 70 % of all samples) instead of in their 8-clock DFMA chains. (One bulk copy
// per lane and chunk was measured too: 96 small cp.async.bulk per warp and chunk serialise in the copy engine, 3x slower.)
// Rows now move as TILES of 32 channels x 16 samples: the warp copies a tile cooperatively with cp.async -- 8 lanes
// cover one 128-byte row segment, so an instruction touches 4 full lines -- into a ring of stages in shared memory,
// several tiles ahead of the arithmetic; results go back the same way (lane rows -> tile -> coalesced 16-byte stores).
// Inside a tile the 16-byte chunk j of row r sits at chunk j ^ (r & 7), which makes both the cooperative accesses and the
// per-lane LDS.128 / STS.128 of a quarter warp conflict-free.
// ------------------------------------------------------------------------------------------
constexpr int kTileT = 16;                        // samples per tile row (128 bytes)
constexpr int kTileD = 32 * kTileT;               // doubles per tile

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// queue the copy of tile k (samples [16k, 16k+16) of rows c0..c0+31 that have their bit set in rowmask); base points at
// sample 0 of row 0 of the array, rows are `stride` doubles apart
__device__ __forceinline__ void tile_fetch(double* tile, const double* base, size_t stride, int c0, int k, uint32_t rowmask, int lane)
{
    const int j = lane & 7;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = 4 * i + (lane >> 3);
        if ((rowmask >> r) & 1u)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(tile + r * kTileT + ((j ^ (r & 7)) << 1))),
                         "l"(base + (size_t)(c0 + r) * stride + (size_t)k * kTileT + 2 * j)
                         : "memory");
    }
}
__device__ __forceinline__ void tile_store(const double* tile, double* base, size_t stride, int c0, int k, uint32_t rowmask, int lane)
{
    // all eight chunks into their own registers first, then the eight stores: a store reads its registers long after it
    // issues, and a load that reuses them has to wait for that (the compiler otherwise cycles through two register quads
    // and the eight load/store pairs serialise on the store pipe's latency)
    const int j = lane & 7;
    double2 v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = 4 * i + (lane >> 3);
        v[i] = *reinterpret_cast<const double2*>(tile + r * kTileT + ((j ^ (r & 7)) << 1));
    }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int r = 4 * i + (lane >> 3);
        if ((rowmask >> r) & 1u) *reinterpret_cast<double2*>(base + (size_t)(c0 + r) * stride + (size_t)k * kTileT + 2 * j) = v[i];
    }
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// the lane's own row of a tile: samples 2j, 2j+1 as one 16-byte access
__device__ __forceinline__ double2 row_ld(const double* tile, int lane, int j)
{
    return *reinterpret_cast<const double2*>(tile + lane * kTileT + ((j ^ (lane & 7)) << 1));
}
__device__ __forceinline__ void row_st(double* tile, int lane, int j, double a, double b)
{
    *reinterpret_cast<double2*>(tile + lane * kTileT + ((j ^ (lane & 7)) << 1)) = make_double2(a, b);
}
// keeps a loop-invariant operand in a register (the compiler otherwise re-reads kernel parameters from the constant bank
// under a predicate inside a select, which turns a two-instruction clamp into a branch)
__device__ __forceinline__ double in_reg(double v) { asm volatile("" : "+d"(v)); return v; }

// A compute warp and its HELPER warp. The compute warp's lanes run one dependent FP64 chain each and should issue nothing
// else: every instruction of an in-order warp that is not part of the chain delays it. So the tile traffic belongs to a
// second warp: the helper queues the cp.async fetches (kStages - 1 tiles ahead; cp.async.mbarrier.arrive.noinc flips the
// tile's `full` barrier when its copies have landed) and writes finished output tiles back; the compute warp only waits
// on `full`, walks its 16 samples out of shared memory, and hands the stage (`empty`) and the output slot (`ofull`) over
// with one elected arrive each. Two output slots alternate (`oempty` returns them).
template <int NIN, int NOUT, int S>
struct TilePipe {
    static constexpr int kBarBytes = 128;
    static constexpr int kDoubles = kBarBytes / 8 + (S * NIN + 2 * NOUT) * kTileD;       // shared memory per warp pair
    uint64_t* bars;
    double* tiles;
    __device__ __forceinline__ TilePipe(double* base) : bars(reinterpret_cast<uint64_t*>(base)), tiles(base + kBarBytes / 8) {}
    __device__ __forceinline__ uint32_t full(int s) const { return smem_u32(bars + s); }
    __device__ __forceinline__ uint32_t empty(int s) const { return smem_u32(bars + S + s); }
    __device__ __forceinline__ uint32_t ofull(int o) const { return smem_u32(bars + 2 * S + o); }
    __device__ __forceinline__ uint32_t oempty(int o) const { return smem_u32(bars + 2 * S + 2 + o); }
    __device__ __forceinline__ double* in_tile(int s, int a) const { return tiles + (s * NIN + a) * kTileD; }
    __device__ __forceinline__ double* out_tile(int o, int a) const { return tiles + (S * NIN + o * NOUT + a) * kTileD; }
    __device__ __forceinline__ void init() const          // one thread of the pair, before the CTA-wide barrier
    {
        for (int s = 0; s < S; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 32;" ::"r"(full(s)));        // every helper lane's copies
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(empty(s)));
        }
        for (int o = 0; o < 2; o++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ofull(o)));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(oempty(o)));
        }
    }
    // try_wait with a suspend-time hint: the waiting warp sleeps in hardware until the phase flips. A compute warp and a
    // helper warp share every SM sub-partition; a warp that polls in a tight loop takes issue slots from the other.
    static __device__ __forceinline__ void wait(uint32_t bar, uint32_t parity)
    {
        asm volatile(
            "{\n\t.reg .pred p;\n"
            "W_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
            "@p bra D_%=;\n\t"
            "bra W_%=;\n"
            "D_%=:\n\t}\n" ::"r"(bar), "r"(parity), "r"(0x989680)
            : "memory");
    }
    static __device__ __forceinline__ void arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
    // whole warp: every lane is past its accesses, then one lane signals
    static __device__ __forceinline__ void warp_arrive(uint32_t bar, int lane)
    {
        __syncwarp();
        if (lane == 0) arrive(bar);
    }
    // ---- compute warp
    __device__ __forceinline__ const double* acquire_in(int k) const { wait(full(k % S), (uint32_t)(k / S) & 1u); return in_tile(k % S, 0); }
    __device__ __forceinline__ void release_in(int k, int lane) const { warp_arrive(empty(k % S), lane); }
    __device__ __forceinline__ double* acquire_out(int k) const
    {
        if (k >= 2) wait(oempty(k & 1), (uint32_t)((k >> 1) - 1) & 1u);
        return out_tile(k & 1, 0);
    }
    __device__ __forceinline__ void release_out(int k, int lane) const { warp_arrive(ofull(k & 1), lane); }
    // ---- helper warp: src[a] / dst[a] point at sample 0 of row 0 of the arrays; in_mask[a] / out_mask[a] pick the rows
    // `post(tile 0 of the slot, chunk)` runs on every helper lane between the compute warp's hand-over and the stores
    template <class Post>
    __device__ __forceinline__ void serve(int nchunks, int c0, int lane, const double* const* src, const size_t* sstride, const uint32_t* in_mask,
                                          double* const* dst, const size_t* dstride, const uint32_t* out_mask, bool with_out, Post&& post) const
    {
        const int iters = with_out ? nchunks + S - 1 : nchunks;
        for (int i = 0; i < iters; i++) {
            if (i < nchunks) {
                const int s = i % S;
                if (i >= S) wait(empty(s), (uint32_t)(i / S - 1) & 1u);
#pragma unroll
                for (int a = 0; a < NIN; a++) tile_fetch(in_tile(s, a), src[a], sstride[a], c0, i, in_mask[a], lane);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full(s)) : "memory");
            }
            const int j = i - (S - 1);
            if (NOUT > 0 && with_out && j >= 0) {
                wait(ofull(j & 1), (uint32_t)(j >> 1) & 1u);
                post(out_tile(j & 1, 0), j);
                __syncwarp();
#pragma unroll
                for (int a = 0; a < NOUT; a++) tile_store(out_tile(j & 1, a), dst[a], dstride[a], c0, j, out_mask[a], lane);
                warp_arrive(oempty(j & 1), lane);
            }
        }
    }
};

// ---- sequential poles. CTA = 4 compute warps + their 4 helper warps; compute warp pairs (2p, 2p+1) run the AGC averagers
// and the S-meter of the same 32 channels. The loops are bound by one dependent FP64 chain per lane, so the warps are
// PACKED on a handful of SMs instead of one 32-thread CTA on each of 64 SMs -- kernel 1T needs a whole SM (every register)
// per CTA and cannot start on an SM that hosts even one of these warps for as long as they live.
constexpr int kSeqPairs = 4;                     // compute warps per CTA (each with a helper warp)
constexpr int kSeqThreads = 64 * kSeqPairs;
constexpr int kSeqStages = 4;
typedef TilePipe<1, 1, kSeqStages> Seq1Pipe;
typedef TilePipe<2, 2, kSeqStages> Seq2Pipe;
constexpr int kSeq1Smem = kSeqPairs * Seq1Pipe::kDoubles * 8;
constexpr int kSeq2Smem = kSeqPairs * Seq2Pipe::kDoubles * 8;
__global__ void __launch_bounds__(kSeqThreads, 1) k_post_seq1(PostBufs b, int n, PostUniform u)
{
    extern __shared__ __align__(128) double sm_tiles[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp % kSeqPairs;
    const bool helper = warp >= kSeqPairs;
    Seq1Pipe pipe(sm_tiles + pair * Seq1Pipe::kDoubles);
    if (!helper && lane == 0) pipe.init();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int gw = blockIdx.x * kSeqPairs + pair;
    const int c0 = (gw >> 1) * 32, c = c0 + lane;
    const bool smeter_warp = (gw & 1) != 0;
    if (c0 >= b.nch) return;
    const bool in_bank = c < b.nch;
    const int mode = in_bank ? b.mode[c] : POST_NONE;
    const int nchunks = n / kTileT, n16 = nchunks * kTileT;
    const bool act = in_bank && (smeter_warp ? mode != POST_AGC_ONLY : PAR(P_AGC_ON) != 0.0);
    const uint32_t mask = __ballot_sync(0xffffffffu, act);
    if (!mask) return;
    if (helper) {
        double* arr = smeter_warp ? b.smag : b.peak;
        const double* src[1] = {arr};
        double* dst[1] = {arr};
        const size_t st[1] = {(size_t)b.row};
        const uint32_t im[1] = {mask}, om[1] = {smeter_warp ? 0u : mask};
        pipe.serve(nchunks, c0, lane, src, st, im, dst, st, om, !smeter_warp, [](double*, int) {});
        return;
    }
    if (smeter_warp) {
        // CSMeter::ProcessData, dsp/smeter.cpp:77-91
        double sm_att = 0, sm_dec = 0, sm_ave = 0, sm_peak = 0;
        if (act) { sm_att = ST(S_SM_ATT); sm_dec = ST(S_SM_DEC); sm_ave = ST(S_SM_AVE); sm_peak = ST(S_SM_PEAK); }
        const double qa = 1.0 - u.sm_attack, qd = 1.0 - u.sm_decay, ka = u.sm_attack, kd = u.sm_decay;
        // Both poles advance with one dependent DFMA each; the reference's "if (att > dec) { ave = att; dec = att; } else
        // ave = dec;" leaves ave == dec == max(att, dec) in either branch, so the step is two FMAs, a compare and a select
        // (no branch: a divergent branch costs several times the 8-clock DFMA latency that bounds this loop).
        auto step = [&](double mag) {
            const double sa = fma(qa, sm_att, ka * mag);
            const double sd = fma(qd, sm_dec, kd * mag);
            sm_att = sa;
            sm_dec = sa > sd ? sa : sd;
            sm_peak = mag > sm_peak ? mag : sm_peak;
        };
        for (int k = 0; k < nchunks; k++) {
            const double* x = pipe.acquire_in(k);
            double2 v[kTileT / 2];
#pragma unroll
            for (int j = 0; j < kTileT / 2; j++) v[j] = row_ld(x, lane, j);
            pipe.release_in(k, lane);
#pragma unroll
            for (int j = 0; j < kTileT / 2; j++) { step(v[j].x); step(v[j].y); }
        }
        if (act) {
            const double* row = b.smag + (size_t)c * b.row;
            for (int t = n16; t < n; t++) step(row[t]);
            if (n > 0) sm_ave = sm_dec;
            ST(S_SM_ATT) = sm_att; ST(S_SM_DEC) = sm_dec; ST(S_SM_AVE) = sm_ave; ST(S_SM_PEAK) = sm_peak;
        }
        return;
    }
    // attack/decay averagers of CAgc::ProcessData, dsp/agc.cpp:235-276; row <- max(attack, decay)
    const int cs = act ? c : c0 + __ffs(mask) - 1;          // idle lanes read a valid channel's parameters
    const bool use_hang = b.par[(size_t)P_AGC_HANG * b.stride + cs] != 0.0;
    const double a_rise = b.par[(size_t)P_A_RISE * b.stride + cs], a_fall = b.par[(size_t)P_A_FALL * b.stride + cs];
    const double d_rise = b.par[(size_t)P_D_RISE * b.stride + cs], d_fall = b.par[(size_t)P_D_FALL * b.stride + cs];
    const int hang_time = (int)b.par[(size_t)P_HANG_TIME * b.stride + cs];
    double att = b.state[(size_t)S_AGC_ATT * b.stride + cs], dec = b.state[(size_t)S_AGC_DEC * b.stride + cs];
    int hang_timer = b.istate[(size_t)I_AGC_HANGT * b.stride + cs];
    // Branch-free step: both candidates of each averager are one DFMA from the state, the comparisons run beside them,
    // the selects pick. Dependent chain per step = DFMA + select instead of compare -> branch -> DMUL -> DFMA.
    const double qar = 1.0 - a_rise, qaf = 1.0 - a_fall, qdr = 1.0 - d_rise, qdf = 1.0 - d_fall;
    auto step = [&](double peak) -> double {
        const double ar = fma(qar, att, a_rise * peak), af = fma(qaf, att, a_fall * peak);
        const double dr = fma(qdr, dec, d_rise * peak), df = fma(qdf, dec, d_fall * peak);
        const bool ga = peak > att, gd = peak > dec;
        const bool hold = use_hang && !gd && hang_timer < hang_time;        // hang: decay frozen while the timer runs
        att = ga ? ar : af;
        dec = gd ? dr : (hold ? dec : df);
        hang_timer = use_hang ? (gd ? 0 : hang_timer + (hold ? 1 : 0)) : hang_timer;
        return att > dec ? att : dec;
    };
    for (int k = 0; k < nchunks; k++) {
        // The warp issues in order: a shared-memory store of step t's result would hold back step t+1 until the select
        // at the end of step t has retired. So the tile's 16 inputs come into registers first, the 16 steps run back to
        // back, and the results leave together at the end (idle lanes run along on harmless values: no divergence).
        const double* x = pipe.acquire_in(k);
        double2 v[kTileT / 2];
#pragma unroll
        for (int j = 0; j < kTileT / 2; j++) v[j] = row_ld(x, lane, j);
        pipe.release_in(k, lane);
#pragma unroll
        for (int j = 0; j < kTileT / 2; j++) { v[j].x = step(v[j].x); v[j].y = step(v[j].y); }
        double* y = pipe.acquire_out(k);
#pragma unroll
        for (int j = 0; j < kTileT / 2; j++) row_st(y, lane, j, v[j].x, v[j].y);
        pipe.release_out(k, lane);
    }
    if (act) {
        double* row = b.peak + (size_t)c * b.row;
        for (int t = n16; t < n; t++) row[t] = step(row[t]);
        ST(S_AGC_ATT) = att; ST(S_AGC_DEC) = dec; IST(I_AGC_HANGT) = hang_timer;
    }
}

// ---- gain law, delayed signal, per-mode pointwise front end; CTA per channel
__global__ void __launch_bounds__(256) k_post_mid(PostBufs b, int n, int delay, int stereo, float* __restrict__ audio,
                                                  int audio_stride, int audio_off, const int* __restrict__ chan_map)
{
    const int c = blockIdx.x;
    const int mode = b.mode[c];
    const bool agc_on = PAR(P_AGC_ON) != 0.0;
    const double knee = PAR(P_KNEE), gain_slope = PAR(P_GAIN_SLOPE), fixed_gain = PAR(P_FIXED_GAIN);
    const double manual_gain = PAR(P_MANUAL_GAIN);
    float2* yrow = b.y + (size_t)c * b.y_row;
    float* aout = audio ? audio + (size_t)chan_map[c] * audio_stride + (stereo ? 2 : 1) * audio_off : nullptr;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        double zr, zi;
        if (agc_on) {
            const double m = b.peak[(size_t)c * b.row + t];
            const double gain = (m <= knee) ? fixed_gain : 0.7 * exp10(m * (gain_slope - 1.0));   // dsp/agc.cpp:278-286 (pow(10., x): exp10 is the same function to an ulp at a third of the instructions)
            const float2 dl = yrow[kYHist + t - delay];            // m_SigDelayBuf: `delay` samples ago
            zr = (double)dl.x * gain;
            zi = (double)dl.y * gain;
        } else {
            const float2 xf = yrow[kYHist + t];
            zr = manual_gain * (double)xf.x;                       // :288-294
            zi = manual_gain * (double)xf.y;
        }
        b.z[(size_t)c * b.row + t] = make_float2((float)zr, (float)zi);
        if (mode == POST_AM) b.u[(size_t)c * b.row + t] = sqrt(zr * zr + zi * zi);          // dsp/amdemod.cpp:72
        else if (mode == POST_SAM) {
            b.u[(size_t)c * b.row + t] = sqrt(zr * zr + zi * zi);
            b.th[(size_t)c * b.row + t] = atan2(zi, zr);
        } else if (mode == POST_FM) b.th[(size_t)c * b.row + t] = atan2(zi, zr);
        else if (mode == POST_SSB) {                                                        // dsp/ssbdemod.cpp:48-60
            if (aout) {
                if (stereo) { aout[2 * t] = (float)zr; aout[2 * t + 1] = (float)zi; }
                else aout[t] = (float)zr;
            }
        }
    }
    __syncthreads();
    // the last kYHist samples of [history | burst] become the next burst's delay history
    float2 keep[kYHist / 256];
#pragma unroll
    for (int q = 0; q < kYHist / 256; q++) keep[q] = yrow[n + threadIdx.x + q * 256];
    __syncthreads();
#pragma unroll
    for (int q = 0; q < kYHist / 256; q++) yrow[threadIdx.x + q * 256] = keep[q];
}

constexpr double kInv2Pi = 1.0 / kTwoPi;
constexpr double kRintMagic = 6755399441055744.0;      // 1.5 * 2^52: (v + magic) - magic == rint(v) for |v| < 2^51
__device__ __forceinline__ double wrap_pi(double d)
{
    return d - kTwoPi * rint(d * (1.0 / kTwoPi));
}

// ---- DC blockers and PLLs; lane per channel, tiles of 32 channels x 16 samples through shared memory; CTA = 4 compute
// warps + their 4 helper warps (see TilePipe)
__global__ void __launch_bounds__(kSeqThreads, 1) k_post_seq2(PostBufs b, int n, PostUniform u, float* __restrict__ audio,
                                                  int audio_stride, int audio_off, const int* __restrict__ chan_map)
{
    extern __shared__ __align__(128) double sm_tiles[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp % kSeqPairs;
    const bool helper = warp >= kSeqPairs;
    Seq2Pipe pipe(sm_tiles + pair * Seq2Pipe::kDoubles);
    if (!helper && lane == 0) pipe.init();
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int c0 = (blockIdx.x * kSeqPairs + pair) * 32, c = c0 + lane;
    if (c0 >= b.nch) return;
    const int mode = c < b.nch ? b.mode[c] : POST_NONE;
    const bool act = mode == POST_AM || mode == POST_SAM || mode == POST_FM;
    const bool sam_st = mode == POST_SAM && u.stereo;
    // rows of the tile arrays this warp pair moves: theta in (SAM, FM), envelope in (AM, SAM), FIR row out (AM, FM),
    // speculative biquad out (FM; it reuses the envelope array, which AM / SAM rows must keep)
    const uint32_t m_th = __ballot_sync(0xffffffffu, mode == POST_SAM || mode == POST_FM);
    const uint32_t m_u = __ballot_sync(0xffffffffu, mode == POST_AM || mode == POST_SAM);
    const uint32_t m_v = __ballot_sync(0xffffffffu, mode == POST_AM || mode == POST_FM);
    const uint32_t m_lp = __ballot_sync(0xffffffffu, mode == POST_FM);
    if (!(m_th | m_u)) return;
    double* vbase = b.v + kHist;
    const int nchunks = n / kTileT, n16 = nchunks * kTileT;
    const bool is_fm = mode == POST_FM;
    const bool tail = n16 < n;
    if (helper) {
        const double* src[2] = {b.th, b.u};
        const size_t sst[2] = {(size_t)b.row, (size_t)b.row};
        const uint32_t im[2] = {m_th, m_u};
        double* dst[2] = {vbase, b.u};
        const size_t dst_st[2] = {(size_t)b.v_row, (size_t)b.row};
        const uint32_t om[2] = {m_v, m_lp};
        // FM: what follows the PLL -- the 10 ms DC tracker on the loop frequency, the output gain and the 3 kHz low-pass
        // biquad (dsp/fmdemod.cpp:184-187, CIir::ProcessFilter dsp/iir.cpp:171-180) -- is a pair of LINEAR recurrences
        // that only consume the PLL's frequency. The compute warp hands the raw frequency over in the output tile; the
        // helper's lanes turn it into the two output rows, so the PLL chain never waits behind them (one in-order warp
        // running both took 180 clocks per sample for a 76-clock chain). The biquad only advances while the squelch is
        // open, which is decided per burst from the whole burst (k_post_fir): it is computed here SPECULATIVELY, outputs
        // to the (otherwise unused) envelope row, end state to S_LP_W1N/W2N; k_post_fir commits or discards them.
        double fm_dc = 0, w1 = 0, w2 = 0;
        if (is_fm) { fm_dc = ST(S_FM_DC); w1 = ST(S_LP_W1); w2 = ST(S_LP_W2); }
        const double qdc = 1.0 - u.fm_dc_alpha, kdc = u.fm_dc_alpha, na1 = -u.lp_a1, na2 = -u.lp_a2;
        auto fm_post = [&](double f, double& lp) -> double {
            fm_dc = fma(qdc, fm_dc, kdc * f);
            const double pre = (f - fm_dc) * u.fm_gain;
            const double w0 = fma(na1, w1, fma(na2, w2, pre));
            lp = fma(u.lp_b0, w0, fma(u.lp_b1, w1, u.lp_b2 * w2));
            w2 = w1;
            w1 = w0;
            return pre;
        };
        pipe.serve(nchunks, c0, lane, src, sst, im, dst, dst_st, om, true, [&](double* o_v, int) {
            if (!is_fm) return;
            double* o_lp = o_v + kTileD;
            double2 f[kTileT / 2], l[kTileT / 2];
#pragma unroll
            for (int j = 0; j < kTileT / 2; j++) f[j] = row_ld(o_v, lane, j);
#pragma unroll
            for (int j = 0; j < kTileT / 2; j++) { f[j].x = fm_post(f[j].x, l[j].x); f[j].y = fm_post(f[j].y, l[j].y); }
#pragma unroll
            for (int j = 0; j < kTileT / 2; j++) { row_st(o_v, lane, j, f[j].x, f[j].y); row_st(o_lp, lane, j, l[j].x, l[j].y); }
        });
        if (tail) {
            // the last n % 16 samples: the compute warp left the raw frequency in the FIR row
            asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");
            if (is_fm) {
                double* vrow = vbase + (size_t)c * b.v_row;
                double* urow = b.u + (size_t)c * b.row;
                for (int t = n16; t < n; t++) { double lp; vrow[t] = fm_post(vrow[t], lp); urow[t] = lp; }
            }
        }
        if (is_fm) { ST(S_FM_DC) = fm_dc; ST(S_LP_W1N) = w1; ST(S_LP_W2N) = w2; }
        return;
    }
    float* aout = (audio && act) ? audio + (size_t)chan_map[c] * audio_stride + (u.stereo ? 2 : 1) * audio_off : nullptr;
    // state (a lane uses the fields of its mode)
    double z1 = 0, y1 = 0, phase = 0, freq = 0;
    if (act) { z1 = ST(S_Z1); y1 = ST(S_Y1); phase = ST(S_PHASE); freq = ST(S_FREQ); }
    // The PLLs (dsp/samdemod.cpp:81-105, dsp/fmdemod.cpp:166-187). SAM: tmp = x e^{-j phase}, err = atan2(tmp) =
    // wrap(arg x - phase); FM: tmp = x e^{+j phase}, err = -atan2(tmp) = -wrap(arg x + phase); arg x comes from k_post_mid.
    // Dependent chain per sample: x (DADD) -> rint through the 1.5 * 2^52 constant (DFMA, DADD; DFRND alone costs 21
    // clocks) -> err (DFMA) -> freq (DFMA) -> clamp (two compare + select; DMNMX costs 25 clocks each) -> phase (DADD).
    // The phase is NOT wrapped per sample: the reference lets it run through the burst and takes fmod once at the end
    // (dsp/fmdemod.cpp:188, dsp/samdemod.cpp:107); err depends on it modulo 2 pi only.
    const double sgn = (is_fm || sam_st) ? 1.0 : -1.0;        // x = th + sgn * phase; err = sgn * (2 pi r - x)
    const double alpha = is_fm ? u.fm_alpha : u.sam_alpha, beta = is_fm ? u.fm_beta : u.sam_beta;
    const double hi = in_reg(is_fm ? u.fm_hi : u.sam_hi), lo = in_reg(is_fm ? u.fm_lo : u.sam_lo);
    const double n2pi = sgn * kTwoPi;
    auto pll = [&](double th) -> double {
        const double x = fma(sgn, phase, th);
        const double r = fma(x, kInv2Pi, kRintMagic) - kRintMagic;
        const double err = fma(r, n2pi, -sgn * x);              // FM / stereo SAM: -(x - 2 pi r); SAM: x - 2 pi r
        const double f = fma(beta, err, freq);
        const double pa = fma(alpha, err, phase);
        const bool over = f > hi, under = f < lo;               // both compares on the unclamped value, side by side
        const double fc = over ? hi : (under ? lo : f);
        freq = fc;
        phase = pa + fc;
        return err;
    };
    // The same step with the NEXT sample's detector input started early: x' = th' + sgn (pa + fc) is formed as
    // fma(sgn, fc, fma(sgn, pa, th')) -- the inner part is ready before the clamp is, which takes the phase addition out
    // of the dependent chain (x -> rint -> err -> freq -> clamp -> x').
    auto pll_x = [&](double x, double th_next) -> double {
        const double r = fma(x, kInv2Pi, kRintMagic) - kRintMagic;
        const double err = fma(r, n2pi, -sgn * x);
        const double f = fma(beta, err, freq);
        const double pa = fma(alpha, err, phase);
        const double t1 = fma(sgn, pa, th_next);
        const bool over = f > hi, under = f < lo;
        const double fc = over ? hi : (under ? lo : f);
        freq = fc;
        phase = pa + fc;
        return fma(sgn, fc, t1);
    };
    // AM: DC removal H(z) = (1 - z^-1)/(1 - .99 z^-1), dsp/amdemod.cpp:73-78
    auto am_step = [&](double mag) -> double {
        const double z0 = fma(z1, 0.99, mag);
        const double o = z0 - z1;
        z1 = z0;
        return o;
    };
    auto sam_step = [&](double th, double mag) -> float {       // tmp.re = |x| cos(err), then the DC blocker
        const double err = pll(th);
        const double z0 = fma(mag, cos(err), z1 * 0.99);
        const float o = (float)(z0 - z1);
        z1 = z0;
        return o;
    };
    // stereo SAM, dsp/samdemod.cpp:115-147: opposite NCO sign to the mono path; BOTH parts of tmp = |x| (cos err, -sin err)
    // are DC-blocked and go on to the Hilbert-pair FIR
    auto sam_stereo_step = [&](double th, double mag, double* o1, double* o2) {
        const double err = pll(th);
        double sn, cs;
        sincos(err, &sn, &cs);
        const double z0 = mag * cs + (z1 * 0.99);
        const double y0 = -mag * sn + (y1 * 0.99);
        *o1 = z0 - z1;
        *o2 = y0 - y1;
        z1 = z0;
        y1 = y0;
    };
    // Every row of the tile is FM (the usual case: a group's channels are sorted by mode): one straight-line loop for
    // the whole warp. The warp issues in order, and ptxas sinks a shared-memory load to just in front of its first use
    // when both sit in one basic block -- a ~30-clock LDS in the dependent chain of every second sample. So the tile's
    // inputs come into registers FIRST, the stage is handed back (the elected arrive ends the basic block), then the 16
    // steps run back to back out of registers; a step's frequency goes to the output tile as soon as it exists (the next
    // step needs it at the same moment, so the store costs an issue slot and no wait).
    if (m_th == m_lp && m_u == 0u) {
        for (int k = 0; k < nchunks; k++) {
            const double* xt = pipe.acquire_in(k);
            double2 v[kTileT / 2];
#pragma unroll
            for (int j = 0; j < kTileT / 2; j++) v[j] = row_ld(xt, lane, j);
            pipe.release_in(k, lane);
            double* o_v = pipe.acquire_out(k);      // the helper turns the loop frequency into the two output rows
            double x = fma(sgn, phase, v[0].x);
#pragma unroll
            for (int j = 0; j < kTileT / 2; j++) {
                x = pll_x(x, v[j].y);
                const double f0 = freq;
                x = pll_x(x, j + 1 < kTileT / 2 ? v[j + 1].x : 0.0);
                row_st(o_v, lane, j, f0, freq);
            }
            pipe.release_out(k, lane);
        }
    } else {
        for (int k = 0; k < nchunks; k++) {
            const double* xt = pipe.acquire_in(k);
            const double* xu = xt + kTileD;
            double* o_v = pipe.acquire_out(k);
            if (is_fm) {
#pragma unroll
                for (int j = 0; j < kTileT / 2; j++) {
                    const double2 v = row_ld(xt, lane, j);
                    pll(v.x);
                    const double f0 = freq;
                    pll(v.y);
                    row_st(o_v, lane, j, f0, freq);
                }
            } else if (mode == POST_AM) {
#pragma unroll
                for (int j = 0; j < kTileT / 2; j++) {
                    const double2 v = row_ld(xu, lane, j);
                    const double p0 = am_step(v.x), p1 = am_step(v.y);
                    row_st(o_v, lane, j, p0, p1);
                }
            } else if (sam_st) {
                double* vrow = vbase + (size_t)c * b.v_row + k * kTileT;
                double* v2row = b.v2 + kHist + (size_t)c * b.v_row + k * kTileT;
                for (int j = 0; j < kTileT / 2; j++) {
                    const double2 t2 = row_ld(xt, lane, j), m2 = row_ld(xu, lane, j);
                    sam_stereo_step(t2.x, m2.x, vrow + 2 * j, v2row + 2 * j);
                    sam_stereo_step(t2.y, m2.y, vrow + 2 * j + 1, v2row + 2 * j + 1);
                }
            } else if (mode == POST_SAM) {
#pragma unroll 2
                for (int j = 0; j < kTileT / 2; j++) {
                    const double2 t2 = row_ld(xt, lane, j), m2 = row_ld(xu, lane, j);
                    const float o0 = sam_step(t2.x, m2.x), o1 = sam_step(t2.y, m2.y);
                    if (aout) { aout[k * kTileT + 2 * j] = o0; aout[k * kTileT + 2 * j + 1] = o1; }
                }
            }
            pipe.release_in(k, lane);
            pipe.release_out(k, lane);
        }
    }
    if (!act) return;
    {   // the last n % 16 samples straight from the rows
        const double* throw_ = b.th + (size_t)c * b.row;
        double* urow = b.u + (size_t)c * b.row;
        double* vrow = vbase + (size_t)c * b.v_row;
        double* v2row = b.v2 + kHist + (size_t)c * b.v_row;
        for (int t = n16; t < n; t++) {
            if (is_fm) { pll(throw_[t]); vrow[t] = freq; }
            else if (mode == POST_AM) vrow[t] = am_step(urow[t]);
            else if (sam_st) sam_stereo_step(throw_[t], urow[t], vrow + t, v2row + t);
            else { const float o = sam_step(throw_[t], urow[t]); if (aout) aout[t] = o; }
        }
    }
    if (tail) asm volatile("bar.sync %0, 64;" ::"r"(1 + pair) : "memory");      // the helper finishes the FM tail
    ST(S_Z1) = z1;
    if (sam_st) ST(S_Y1) = y1;
    if (mode != POST_AM) { ST(S_PHASE) = wrap_pi(phase); ST(S_FREQ) = freq; }    // fmod(m_NcoPhase, K_2PI) at the end of the burst
}

// ---- Kaiser FIRs, squelch, biquad; CTA per channel. dynamic smem: (kHist + max_n + kFirMax) doubles
__global__ void __launch_bounds__(256) k_post_fir(PostBufs b, int n, PostUniform u, float* __restrict__ audio, int audio_stride,
                                                  int audio_off, const int* __restrict__ chan_map)
{
    extern __shared__ double sm_d[];
    __shared__ double red[8];
    __shared__ int s_squelched;
    const int c = blockIdx.x;
    const int mode = b.mode[c];
    const int stereo = u.stereo;
    float* aout = audio ? audio + (size_t)chan_map[c] * audio_stride + (stereo ? 2 : 1) * audio_off : nullptr;
    if (mode == POST_SAM && stereo) {
        // Hilbert-pair band-pass on (re, im) with separate real coefficient sets, then the sideband
        // split L = re + im (lower), R = re - im (upper): dsp/samdemod.cpp:148-156, dsp/fir.cpp:101-127
        double* vre = sm_d;
        double* vim = sm_d + kHist + b.row;
        __shared__ double hi[kFirMax], hq[kFirMax];
        double* r1 = b.v + (size_t)c * b.v_row;
        double* r2 = b.v2 + (size_t)c * b.v_row;
        for (int i = threadIdx.x; i < kHist + n; i += blockDim.x) { vre[i] = r1[i]; vim[i] = r2[i]; }
        for (int i = threadIdx.x; i < kFirMax; i += blockDim.x) { hi[i] = b.sam_taps[i]; hq[i] = b.sam_taps[kFirMax + i]; }
        __syncthreads();
        for (int i = threadIdx.x; i < kHist; i += blockDim.x) { r1[i] = vre[n + i]; r2[i] = vim[n + i]; }
        for (int t = threadIdx.x; t < n; t += blockDim.x) {
            double are = 0.0, aim = 0.0;
            for (int k = 0; k < u.sam_ntaps; k++) { are += hi[k] * vre[kHist + t - k]; aim += hq[k] * vim[kHist + t - k]; }
            if (aout) { aout[2 * t] = (float)(are + aim); aout[2 * t + 1] = (float)(are - aim); }
        }
        return;
    }
    if (mode != POST_AM && mode != POST_FM) return;
    // The input row (kHist samples of history + n new ones) is staged in shared memory as four residue planes,
    // sample p at plane (p & 3), position (p >> 2). A thread computes FOUR consecutive outputs with a sliding
    // register window, so every tap costs one conflict-free shared load of x (consecutive positions across the
    // lanes) and one broadcast load of the coefficient per four multiply-adds -- a quarter of the shared-memory
    // traffic of one load per product. Each output still sums its products in ascending tap order.
    // AM: the filter output IS the audio -> double. FM: the filter (squelch high-pass) only feeds a rectified average that
    // is compared with a threshold in steps of 100 -> float32 operands (half the shared-memory bytes, twice the FMA rate).
    const int Q = ((kHist + n + 3) >> 2) + 1;           // plane stride (elements)
    const bool f32 = mode == POST_FM;
    double* v = sm_d;                                   // 4 * Q doubles  (<= kHist + row + 8)
    double* h = sm_d + kHist + b.row + 8;
    float* vf = reinterpret_cast<float*>(sm_d);
    float* hf = reinterpret_cast<float*>(sm_d + kHist + b.row + 8);
    const int ntaps = (int)PAR(P_NTAPS);
    double* vrow = b.v + (size_t)c * b.v_row;
    if (f32) {
        for (int i = threadIdx.x; i < kHist + n; i += blockDim.x) vf[(i & 3) * Q + (i >> 2)] = (float)vrow[i];
        for (int i = threadIdx.x; i < kFirMax; i += blockDim.x) hf[i] = (float)b.taps[(size_t)c * kFirMax + i];
    } else {
        for (int i = threadIdx.x; i < kHist + n; i += blockDim.x) v[(i & 3) * Q + (i >> 2)] = vrow[i];
        for (int i = threadIdx.x; i < kFirMax; i += blockDim.x) h[i] = b.taps[(size_t)c * kFirMax + i];
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kHist; i += blockDim.x) {       // FIR history for the next burst
        const int p = n + i;
        vrow[i] = f32 ? (double)vf[(p & 3) * Q + (p >> 2)] : v[(p & 3) * Q + (p >> 2)];
    }
    // acc[j] = sum_k h[k] x[kHist + 4g + j - k]   (CFir::ProcessFilter, dsp/fir.cpp:72-91)
    auto fir4_t = [&](int g, auto* acc, const auto* vv, const auto* hh) {
        typedef typename std::remove_const<typename std::remove_pointer<decltype(vv)>::type>::type T;
        const int p0 = kHist + 4 * g;
        T w0 = vv[((p0)&3) * Q + ((p0) >> 2)], w1 = vv[((p0 + 1) & 3) * Q + ((p0 + 1) >> 2)],
          w2 = vv[((p0 + 2) & 3) * Q + ((p0 + 2) >> 2)], w3 = vv[((p0 + 3) & 3) * Q + ((p0 + 3) >> 2)];
        acc[0] = acc[1] = acc[2] = acc[3] = (T)0;
        for (int k = 0; k < ntaps; k++) {
            const T hk = hh[k];
            acc[0] += hk * w0;
            acc[1] += hk * w1;
            acc[2] += hk * w2;
            acc[3] += hk * w3;
            const int p = p0 - (k + 1);
            w3 = w2; w2 = w1; w1 = w0;
            w0 = p >= 0 ? vv[(p & 3) * Q + (p >> 2)] : (T)0;
        }
    };
    auto fir4 = [&](int g, double* acc) { fir4_t(g, acc, (const double*)v, (const double*)h); };
    if (mode == POST_AM) {
        // post filter of dsp/amdemod.cpp:80
        for (int g = threadIdx.x; 4 * g < n; g += blockDim.x) {
            double acc[4];
            fir4(g, acc);
            if (aout) {
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int t = 4 * g + j;
                    if (t >= n) break;
                    if (stereo) { aout[2 * t] = (float)acc[j]; aout[2 * t + 1] = (float)acc[j]; }   // dsp/amdemod.cpp:87-104
                    else aout[t] = (float)acc[j];
                }
            }
        }
        return;
    }
    // FM: PerformNoiseSquelch, dsp/fmdemod.cpp:113-152. The squelch average is a one-pole over
    // |high-passed audio|; only its value at the END of the burst is used, which is the weighted sum
    //   (1-a)^n s0 + a * sum_t (1-a)^(n-1-t) |hp[t]|
    double part = 0.0;
    for (int g = threadIdx.x; 4 * g < n; g += blockDim.x) {
        float acc[4];
        fir4_t(g, acc, (const float*)vf, (const float*)hf);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int t = 4 * g + j;
            if (t < n) part += (double)fabsf(acc[j]) * b.qpow[n - 1 - t];
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) part += __shfl_xor_sync(0xffffffffu, part, d);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x == 0) {
        double sum = 0.0;
        for (int w = 0; w < 8; w++) sum += red[w];
        const double sq_ave = b.qpow[n] * ST(S_SQ_AVE) + u.fm_sq_alpha * sum;
        ST(S_SQ_AVE) = sq_ave;
        int squelched = IST(I_SQUELCHED);
        const double sq_thresh = PAR(P_SQ_THRESH);
        if (0 == sq_thresh) squelched = 1;
        else if (squelched) { if (sq_ave < (sq_thresh - 100.0)) squelched = 0; }
        else { if (sq_ave >= (sq_thresh + 100.0)) squelched = 1; }
        IST(I_SQUELCHED) = squelched;
        s_squelched = squelched;
        if (!squelched) { ST(S_LP_W1) = ST(S_LP_W1N); ST(S_LP_W2) = ST(S_LP_W2N); }   // commit the speculative biquad state
    }
    __syncthreads();
    if (aout) {
        const double* lprow = b.u + (size_t)c * b.row;
        // stereo FM copies the mono stream into both channels (dsp/fmdemod.cpp:229-234)
        const int m = stereo ? 2 * n : n;
        if (s_squelched) for (int t = threadIdx.x; t < m; t += blockDim.x) aout[t] = 0.f;
        else for (int t = threadIdx.x; t < m; t += blockDim.x) aout[t] = (float)lprow[stereo ? (t >> 1) : t];
    }
}
#undef PAR
#undef ST
#undef IST

// ------------------------------------------------------------------------------------------
// PostBank
// ------------------------------------------------------------------------------------------
PostBank::~PostBank()
{
    cudaFree(d_par_); cudaFree(d_taps_); cudaFree(d_mode_); cudaFree(d_reset_); cudaFree(d_state_);
    cudaFree(d_istate_); cudaFree(d_y_); cudaFree(d_magh_); cudaFree(d_smag_); cudaFree(d_peak_); cudaFree(d_z_);
    cudaFree(d_u_); cudaFree(d_th_); cudaFree(d_v_); cudaFree(d_v2_); cudaFree(d_sam_taps_); cudaFree(d_qpow_);
}

int PostBank::init(int nch, int stride, double rate, int max_samples, cudaStream_t st, LaunchCounter* lc)
{
    nch_ = nch; stride_ = stride; rate_ = rate; st_ = st; lc_ = lc;
    max_n_ = round_up(max_samples, 2);      // rows stay 16-byte aligned for the bulk copies of the sequential kernels
    y_row_ = kYHist + max_n_;
    v_row_ = round_up(kHist + max_n_, 2);
    // CSMeter, dsp/smeter.cpp:66-69
    uni_.sm_attack = (1.0 - exp(-1.0 / (rate * .01)));
    uni_.sm_decay = (1.0 - exp(-1.0 / (rate * .5)));
    // CAgc, dsp/agc.cpp:157-164
    uni_.agc_delay = (int)(rate * .015);
    uni_.agc_window = (int)(rate * .018);
    if (uni_.agc_delay >= kAgcBuf - 1) uni_.agc_delay = kAgcBuf - 1;
    if (uni_.agc_window > kAgcBuf) { set_error("AGC window %d exceeds MAX_DELAY_BUF at %g Hz", uni_.agc_window, rate); return CUTESDR_E_ARG; }
    if (uni_.agc_delay < 1 || uni_.agc_window < 1) { set_error("sample rate %g too low for the AGC", rate); return CUTESDR_E_ARG; }
    const double norm = kTwoPi / rate;
    // CSamDemod ctor, dsp/samdemod.cpp:59-65
    uni_.sam_lo = -1000.0 * norm; uni_.sam_hi = 1000.0 * norm;
    uni_.sam_alpha = 2.0 * .707 * 100.0 * norm;
    uni_.sam_beta = (uni_.sam_alpha * uni_.sam_alpha) / (4.0 * .707 * .707);
    // CFmDemod ctor, dsp/fmdemod.cpp:68-86
    uni_.fm_lo = -6000.0 * norm; uni_.fm_hi = 6000.0 * norm;
    uni_.fm_alpha = 2.0 * .707 * 3000.0 * 2.0 * norm;
    uni_.fm_beta = (uni_.fm_alpha * uni_.fm_alpha) / (4.0 * .707 * .707);
    uni_.fm_gain = 25000.0 / uni_.fm_hi;
    uni_.fm_dc_alpha = (1.0 - exp(-1.0 / (rate * 0.01)));
    uni_.fm_sq_alpha = (1.0 - exp(-1.0 / (rate * .02)));
    {   // CIir::InitLP(3000, 1.0, rate), dsp/iir.cpp:86-101
        const double w0 = kTwoPi * 3000.0 / rate, alpha = sin(w0) / 2.0, A = 1.0 / (1.0 + alpha);
        uni_.lp_b0 = A * ((1.0 - cos(w0)) / 2.0);
        uni_.lp_b1 = A * (1.0 - cos(w0));
        uni_.lp_b2 = A * ((1.0 - cos(w0)) / 2.0);
        uni_.lp_a1 = A * (-2.0 * cos(w0));
        uni_.lp_a2 = A * (1.0 - alpha);
    }
    agc_.assign(stride, AgcHost());
    fm_bw_.assign(stride, 3000.0);
    h_par_.assign((size_t)P_COUNT * stride, 0.0);
    h_taps_.assign((size_t)kFirMax * stride, 0.0);
    h_mode_.assign(stride, POST_NONE);
    h_reset_.assign(stride, 0);
    for (int i = 0; i < stride; i++) {
        h_reset_[i] = R_AGC | R_DEMOD | R_FIR | R_SMETER;
        h_par_[(size_t)P_AGC_ON * stride + i] = 1.0;
        h_par_[(size_t)P_NTAPS * stride + i] = 1.0;
    }
    const size_t rows = (size_t)stride;
    CSDR_CK(cudaMalloc(&d_par_, h_par_.size() * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_taps_, h_taps_.size() * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_mode_, stride * sizeof(int)));
    CSDR_CK(cudaMalloc(&d_reset_, stride * sizeof(int)));
    CSDR_CK(cudaMalloc(&d_state_, (size_t)S_COUNT * stride * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_istate_, (size_t)I_COUNT * stride * sizeof(int)));
    CSDR_CK(cudaMalloc(&d_y_, rows * y_row_ * sizeof(float2)));
    CSDR_CK(cudaMalloc(&d_magh_, rows * kAgcBuf * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_smag_, rows * max_n_ * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_peak_, rows * max_n_ * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_z_, rows * max_n_ * sizeof(float2)));
    CSDR_CK(cudaMalloc(&d_u_, rows * max_n_ * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_th_, rows * max_n_ * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_v_, rows * v_row_ * sizeof(double)));
    CSDR_CK(cudaMalloc(&d_v2_, rows * v_row_ * sizeof(double)));
    CSDR_CK(cudaMemsetAsync(d_v2_, 0, rows * v_row_ * sizeof(double), st_));
    {   // CSamDemod ctor, dsp/samdemod.cpp:67-72: 40 dB Kaiser LP 4500/5500 Hz turned into a Hilbert pair at 5 kHz
        double coef[kFirMax];
        std::vector<double> iq(2 * kFirMax, 0.0);
        const int nt = design_kaiser_lp(1.0, 40.0, 4500, 5500, rate, coef);
        for (int k = 0; k < nt; k++) {      // CFir::GenerateHBFilter, dsp/fir.cpp:374-386
            const double arg = (kTwoPi * 5000.0 / rate) * ((double)k - ((double)(nt - 1) / 2.0));
            iq[k] = 2.0 * coef[k] * cos(arg);
            iq[kFirMax + k] = 2.0 * coef[k] * sin(arg);
        }
        uni_.sam_ntaps = nt;
        CSDR_CK(cudaMalloc(&d_sam_taps_, iq.size() * sizeof(double)));
        CSDR_CK(cudaMemcpy(d_sam_taps_, iq.data(), iq.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    {
        std::vector<double> qp(max_n_ + 1);
        const double q = 1.0 - uni_.fm_sq_alpha;
        qp[0] = 1.0;
        for (int j = 1; j <= max_n_; j++) qp[j] = qp[j - 1] * q;
        CSDR_CK(cudaMalloc(&d_qpow_, qp.size() * sizeof(double)));
        CSDR_CK(cudaMemcpy(d_qpow_, qp.data(), qp.size() * sizeof(double), cudaMemcpyHostToDevice));
    }
    CSDR_CK(cudaMemsetAsync(d_state_, 0, (size_t)S_COUNT * stride * sizeof(double), st_));
    CSDR_CK(cudaMemsetAsync(d_istate_, 0, (size_t)I_COUNT * stride * sizeof(int), st_));
    CSDR_CK(cudaMemsetAsync(d_y_, 0, rows * y_row_ * sizeof(float2), st_));
    CSDR_CK(cudaMemsetAsync(d_v_, 0, rows * v_row_ * sizeof(double), st_));
    const size_t smem_pre = 2 * (size_t)(kAgcBuf + max_n_) * sizeof(double);
    const size_t smem_fir = (size_t)(2 * (kHist + max_n_) + kFirMax + 16) * sizeof(double);
    if (smem_pre > 200 * 1024) { set_error("post: burst capacity %d too large", max_n_); return CUTESDR_E_ARG; }
    CSDR_CK(cudaFuncSetAttribute(k_post_pre, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pre));
    CSDR_CK(cudaFuncSetAttribute(k_post_seq1, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeq1Smem));
    CSDR_CK(cudaFuncSetAttribute(k_post_seq2, cudaFuncAttributeMaxDynamicSharedMemorySize, kSeq2Smem));
    CSDR_CK(cudaFuncSetAttribute(k_post_fir, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem_fir, 1024)));
    dirty_ = true;
    return CUTESDR_OK;
}

void PostBank::set_mode(int i, int mode)
{
    int m;
    switch (mode) {
    case CUTESDR_DEMOD_AM: m = POST_AM; break;
    case CUTESDR_DEMOD_SAM: m = POST_SAM; break;
    case CUTESDR_DEMOD_FM: m = POST_FM; break;
    case CUTESDR_DEMOD_USB: case CUTESDR_DEMOD_LSB: case CUTESDR_DEMOD_CWU: case CUTESDR_DEMOD_CWL: m = POST_SSB; break;
    case POST_AGC_ONLY: m = POST_AGC_ONLY; break;
    default: m = POST_NONE; break;
    }
    h_mode_[i] = m;
    h_reset_[i] |= R_DEMOD | R_FIR;
    if (m == POST_FM) {   // CFmDemod ctor designs its squelch high-pass for 3 kHz (dsp/fmdemod.cpp:79,88)
        fm_bw_[i] = -1.0;
    }
    dirty_ = true;
}

void PostBank::free_channel(int i)
{
    // the slot's next user is a new CDemodulator: new CAgc, CSMeter and demodulator objects
    h_mode_[i] = POST_NONE;
    h_reset_[i] = R_AGC | R_DEMOD | R_FIR | R_SMETER;
    agc_[i] = AgcHost();
    fm_bw_[i] = 3000.0;
    h_par_[(size_t)P_AGC_ON * stride_ + i] = 1.0;
    h_par_[(size_t)P_NTAPS * stride_ + i] = 1.0;
    dirty_ = true;
}

void PostBank::set_agc(int i, int on, int hang, int thresh, int manual_gain, int slope, int decay)
{
    // CAgc::SetParameters, dsp/agc.cpp:104-167 (the rate never changes inside a group)
    AgcHost& a = agc_[i];
    if (a.valid && on == a.on && hang == a.hang && thresh == a.thresh && manual_gain == a.mgain && slope == a.slope &&
        decay == a.decay)
        return;
    a.valid = true; a.on = on; a.hang = hang; a.thresh = thresh; a.mgain = manual_gain; a.slope = slope; a.decay = decay;
    auto P = [&](int f) -> double& { return h_par_[(size_t)f * stride_ + i]; };
    P(P_AGC_ON) = on ? 1.0 : 0.0;
    P(P_AGC_HANG) = hang ? 1.0 : 0.0;
    P(P_MANUAL_GAIN) = 32767.0 * pow(10.0, -(100 - (double)manual_gain) / 20.0);
    const double knee = (double)thresh / 20.0;
    const double gain_slope = a.slope / (100.0);
    P(P_KNEE) = knee;
    P(P_GAIN_SLOPE) = gain_slope;
    P(P_FIXED_GAIN) = 0.7 * pow(10.0, knee * (gain_slope - 1.0));
    P(P_A_RISE) = (1.0 - exp(-1.0 / (rate_ * .002)));
    P(P_A_FALL) = (1.0 - exp(-1.0 / (rate_ * .005)));
    P(P_D_RISE) = (1.0 - exp(-1.0 / (rate_ * (double)decay * .001 * .3)));
    P(P_HANG_TIME) = (double)(int)(rate_ * (double)decay * .001);
    if (hang) P(P_D_FALL) = (1.0 - exp(-1.0 / (rate_ * .05)));
    else P(P_D_FALL) = (1.0 - exp(-1.0 / (rate_ * (double)decay * .001)));
    dirty_ = true;
}

void PostBank::set_am_bandwidth(int i, double bw)
{
    double coef[kFirMax];
    int n = design_kaiser_lp(1.0, 50.0, bw, bw * 1.8, rate_, coef);
    for (int k = 0; k < kFirMax; k++) h_taps_[(size_t)i * kFirMax + k] = k < n ? coef[k] : 0.0;
    h_par_[(size_t)P_NTAPS * stride_ + i] = n;
    h_reset_[i] |= R_FIR;
    dirty_ = true;
    taps_dirty_ = true;
}

void PostBank::set_fm(int i, int squelch_value, double fm_bw)
{
    h_par_[(size_t)P_SQ_THRESH * stride_ + i] = (double)(5000.0 - ((5000.0 * squelch_value) / 99));
    if (fm_bw_[i] != fm_bw) {     // dsp/fmdemod.cpp:160-164 -> InitNoiseSquelch re-designs and clears the HP FIR
        fm_bw_[i] = fm_bw;
        double coef[kFirMax];
        int n = design_kaiser_hp(1.0, 50.0, fm_bw, fm_bw * .6, rate_, coef);
        for (int k = 0; k < kFirMax; k++) h_taps_[(size_t)i * kFirMax + k] = k < n ? coef[k] : 0.0;
        h_par_[(size_t)P_NTAPS * stride_ + i] = n;
        h_reset_[i] |= R_FIR;
        taps_dirty_ = true;
    }
    dirty_ = true;
}

int PostBank::upload()
{
    if (!dirty_) return CUTESDR_OK;
    // snapshots go through pinned staging buffers: no stream synchronisation, and the host vectors may be edited
    // again right away
    CSDR_TRY(stage_.upload(d_par_, h_par_.data(), h_par_.size() * sizeof(double), st_));
    if (taps_dirty_) CSDR_TRY(stage_.upload(d_taps_, h_taps_.data(), h_taps_.size() * sizeof(double), st_));
    taps_dirty_ = false;
    CSDR_TRY(stage_.upload(d_mode_, h_mode_.data(), stride_ * sizeof(int), st_));
    bool any_reset = false;
    for (int r : h_reset_) any_reset |= (r != 0);
    if (any_reset) CSDR_TRY(stage_.upload(d_reset_, h_reset_.data(), stride_ * sizeof(int), st_));
    if (any_reset) {
        std::fill(h_reset_.begin(), h_reset_.end(), 0);
        need_reset_kernel_ = true;
    }
    dirty_ = false;
    return CUTESDR_OK;
}

int PostBank::run(int n, float* d_audio, int audio_stride, int audio_off, const int* d_chan_map)
{
    if (n <= 0) return CUTESDR_OK;
    if (n > max_n_) { set_error("PostBank::run: %d samples exceed capacity %d", n, max_n_); return CUTESDR_E_ARG; }
    CSDR_TRY(upload());
    PostBufs b;
    b.y = d_y_; b.y_row = y_row_; b.magh = d_magh_; b.smag = d_smag_; b.peak = d_peak_; b.z = d_z_; b.u = d_u_; b.th = d_th_;
    b.row = max_n_; b.v = d_v_; b.v2 = d_v2_; b.sam_taps = d_sam_taps_; b.v_row = v_row_; b.par = d_par_; b.taps = d_taps_; b.mode = d_mode_; b.reset = d_reset_;
    b.state = d_state_; b.istate = d_istate_; b.nch = nch_; b.stride = stride_; b.qpow = d_qpow_;
    if (need_reset_kernel_) {
        k_post_reset<<<stride_, 128, 0, st_>>>(b);      // every slot: parked slots keep their flags until used
        lc_->n++;
        need_reset_kernel_ = false;
    }
    const size_t smem_fir = (size_t)((uni_.stereo ? 2 : 1) * (kHist + max_n_) + kFirMax + 16) * sizeof(double);
    const int seq_blocks = (nch_ + 31) / 32;
    k_post_pre<<<nch_, 256, 2 * (size_t)(uni_.agc_window - 1 + n) * sizeof(double), st_>>>(b, n, uni_.agc_window);
    k_post_seq1<<<(2 * seq_blocks + kSeqPairs - 1) / kSeqPairs, kSeqThreads, kSeq1Smem, st_>>>(b, n, uni_);
    k_post_mid<<<nch_, 256, 0, st_>>>(b, n, uni_.agc_delay, uni_.stereo, d_audio, audio_stride, audio_off, d_chan_map);
    k_post_seq2<<<(seq_blocks + kSeqPairs - 1) / kSeqPairs, kSeqThreads, kSeq2Smem, st_>>>(b, n, uni_, d_audio, audio_stride, audio_off, d_chan_map);
    k_post_fir<<<nch_, 256, smem_fir, st_>>>(b, n, uni_, d_audio, audio_stride, audio_off, d_chan_map);
    lc_->n += 5;
    CSDR_CK(cudaGetLastError());
    return CUTESDR_OK;
}

int PostBank::read_smeter(int i, double* peak, double* ave)
{
    double pk = 0, av = 0;
    CSDR_CK(cudaMemcpyAsync(&pk, d_state_ + (size_t)S_SM_PEAK * stride_ + i, sizeof(double), cudaMemcpyDeviceToHost, st_));
    CSDR_CK(cudaMemcpyAsync(&av, d_state_ + (size_t)S_SM_AVE * stride_ + i, sizeof(double), cudaMemcpyDeviceToHost, st_));
    // GetPeak resets the held peak (dsp/smeter.cpp:99-104); GetAve does not (:109-112), so the reset only happens
    // when the caller asks for the peak
    if (peak) CSDR_CK(cudaMemsetAsync(d_state_ + (size_t)S_SM_PEAK * stride_ + i, 0, sizeof(double), st_));
    CSDR_CK(cudaStreamSynchronize(st_));
    if (peak) *peak = pk + 5.0;     // SMETER_CALIBRATION, dsp/smeter.cpp:45
    if (ave) *ave = av + 5.0;
    return CUTESDR_OK;
}

}  // namespace csdr
