// bank.cu -- receiver bank (N x CDemodulator) and its C ABI.
// Sequencing mirrors CDemodulator::SetDemod / ProcessData, dsp/demodulator.cpp:107-215.
#include "bank.cuh"
#include "mgpu.cuh"

#include <algorithm>
#include <stdlib.h>

namespace csdr {

static thread_local std::string g_err;

void set_error(const char* fmt, ...)
{
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

Group::~Group()
{
    if (st_post) { cudaStreamSynchronize(st_post); }
    if (ev_dec) cudaEventDestroy(ev_dec);
    for (auto& e : ev_post) if (e) cudaEventDestroy(e);
    cudaFree(d_demod);
    cudaFree(d_chan_map);
    cudaFree(d_local_map);
    if (st_post) cudaStreamDestroy(st_post);
}

TapSpectrum::~TapSpectrum()
{
    cudaFree(d_frame);
    for (auto& e : ev) if (e) cudaEventDestroy(e);
}

// kind 0: complex rows (float2); 1: real samples -> (x, 0) (TYPEREAL DisplayData, gui/testbench.cpp:657-658)
__global__ void k_tap_append(float2* __restrict__ dst, const void* __restrict__ src, int n, int kind)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    dst[i] = kind ? make_float2(reinterpret_cast<const float*>(src)[i], 0.f) : reinterpret_cast<const float2*>(src)[i];
}

static int block_limit(double in_rate, double out_rate)
{
    // m_InBufLimit, dsp/demodulator.cpp:145-146 (same expression, same evaluation order)
    int limit = (out_rate / 100.0) * in_rate / out_rate;
    limit &= 0xFFFFFF00;
    return limit;
}

}  // namespace csdr

using namespace csdr;

cutesdr_bank::~cutesdr_bank()
{
    if (st) cudaStreamSynchronize(st);
    groups.clear();
    nb.reset();
    cudaFree(d_x);
    for (int k = 0; k < kAsyncSlots; k++) {
        cudaFree(d_xs[k]);
        if (ev_h2d[k]) cudaEventDestroy(ev_h2d[k]);
        if (ev_free[k]) cudaEventDestroy(ev_free[k]);
        if (ev_host[k]) cudaEventDestroy(ev_host[k]);
    }
    if (ev_d2h) cudaEventDestroy(ev_d2h);
    if (st_h2d) { cudaStreamSynchronize(st_h2d); cudaStreamDestroy(st_h2d); }
    if (st_d2h) { cudaStreamSynchronize(st_d2h); cudaStreamDestroy(st_d2h); }
    cudaFree(d_halo[0]);
    cudaFree(d_halo[1]);
    cudaFree(d_audio);
    if (h_stage) cudaFreeHost(h_stage);
    if (st) cudaStreamDestroy(st);
}

// push the per-channel parameters of user channel c into its group's stage objects
static int apply_channel(cutesdr_bank* b, int c, bool new_demod)
{
    ChanCfg& cc = b->ch[c];
    Group& g = *b->groups[cc.group];
    const int i = cc.local;
    g.dec.set_frequency(i, cc.dc_nco_freq);
    CSDR_TRY(g.fir.setup(i, cc.info.LowCut, cc.info.HiCut, cc.demod_cw, g.dec.out_rate()));
    if (new_demod) g.post.set_mode(i, cc.mode);
    g.post.set_agc(i, cc.info.AgcOn, cc.info.AgcHangOn, cc.info.AgcThresh, cc.info.AgcManualGain, cc.info.AgcSlope,
                   cc.info.AgcDecay);
    if (cc.mode == CUTESDR_DEMOD_FM) g.post.set_fm(i, cc.info.SquelchValue, (double)cc.info.HiCut);
    if (cc.mode == CUTESDR_DEMOD_AM) g.post.set_am_bandwidth(i, (cc.info.HiCut - cc.info.LowCut) / 2.0);
    return CUTESDR_OK;
}

// A group of channels that share the decimation chain for max_bw: its stage objects, burst stream and maps.
// min_cap > number of channels leaves parked slots for channels that join later (move_channel).
int cutesdr_bank::create_group(double max_bw, const std::vector<int>& chans, int min_cap, int* index)
{
    std::unique_ptr<Group> g(new Group());
    g->max_bw = max_bw;
    g->chans = chans;
    const int n = (int)g->chans.size();
    {   // burst kernels get scheduled ahead of kernel 1's queued CTAs whenever an SM slot frees up
        int lo = 0, hi = 0;
        CSDR_CK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        CSDR_CK(cudaStreamCreateWithPriority(&g->st_post, cudaStreamNonBlocking, hi));
    }
    CSDR_CK(cudaEventCreateWithFlags(&g->ev_dec, cudaEventDisableTiming));
    for (auto& e : g->ev_post) CSDR_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CSDR_TRY(g->dec.init(std::max(n, min_cap), in_rate, g->max_bw, L, st, &lc));
    CSDR_TRY(g->dec.set_overlap(getenv("CUTESDR_OVERLAP_K2") != nullptr));   // measured: no gain
    const int stride = g->dec.stride();
    g->cap = stride;
    CSDR_TRY(g->fir.init(n, stride, g->st_post, &lc));
    CSDR_TRY(g->post.init(n, stride, g->dec.out_rate(), kMaxBurstSamples, g->st_post, &lc));
    g->post.set_stereo(stereo);
    CSDR_CK(cudaMalloc(&g->d_chan_map, stride * sizeof(int)));
    CSDR_CK(cudaMalloc(&g->d_local_map, stride * sizeof(int)));
    std::vector<int> ident(stride, 0);
    g->h_chan_map.assign(stride, 0);
    for (int i = 0; i < stride; i++) ident[i] = i;
    for (int i = 0; i < n; i++) g->h_chan_map[i] = g->chans[i];
    CSDR_CK(cudaMemcpy(g->d_chan_map, g->h_chan_map.data(), stride * sizeof(int), cudaMemcpyHostToDevice));
    CSDR_CK(cudaMemcpy(g->d_local_map, ident.data(), stride * sizeof(int), cudaMemcpyHostToDevice));
    if (audio_rate > 0.0) {
        g->rs.reset(new ResamplerBank());
        // stereo: rows of (left,right) pairs through the TYPECPX form of the resampler (dsp/fractresampler.cpp:194-249)
        CSDR_TRY(g->rs->init(stride, kMaxBurstSamples, g->st_post, &lc, stereo ? 2 : 1));
        g->rs->set_rows(n);
    }
    const int gi = (int)groups.size();
    for (int i = 0; i < n; i++) { ch[g->chans[i]].group = gi; ch[g->chans[i]].local = i; }
    groups.push_back(std::move(g));
    if (index) *index = gi;
    return CUTESDR_OK;
}

// CDemodulator::SetDemod with a new m_DesiredMaxOutputBandwidth on a RUNNING bank (dsp/demodulator.cpp:111-142): the
// reference rebuilds that one object's decimation chain and demodulator and leaves every other CDemodulator alone. Here
// the channel leaves its group's slot and takes a parked slot of a group with the new chain (or a new group); only its
// own rows are re-initialised, in stream order. The other channels' state, the wideband halo and the stream position
// are untouched -- their output is bit-identical to a run without the change.
int cutesdr_bank::move_channel(int c)
{
    ChanCfg& cc = ch[c];
    std::vector<int> lens;
    const double orate = plan_stages(in_rate, cc.max_bw, lens);
    if (block_limit(in_rate, orate) != block_limit(in_rate, groups[cc.group]->dec.out_rate()) || L % (1 << lens.size()) != 0) {
        layout_dirty = true;          // the new chain needs another DSP block length: only a full rebuild can do that
        return CUTESDR_OK;
    }
    {   // leave the old slot
        const int ga = cc.group;
        Group& a = *groups[ga];
        a.chans[cc.local] = -1;
        a.fir.release(cc.local);
        a.post.free_channel(cc.local);
        while (!a.chans.empty() && a.chans.back() < 0) a.chans.pop_back();
        const int used = (int)a.chans.size();
        a.fir.set_nch(used);
        a.post.set_nch(used);
        if (a.rs) a.rs->set_rows(used);
        if (used == 0) {
            // last channel gone: drop the group (its destructor waits for its own burst stream only)
            CSDR_TRY(join());
            CSDR_CK(cudaStreamSynchronize(st));
            groups.erase(groups.begin() + ga);
            for (auto& o : ch) if (o.group > ga) o.group--;
        }
        cc.group = -1;
        cc.local = -1;
    }
    int gb = -1, slot = -1, biggest = 0;
    for (size_t gi = 0; gi < groups.size() && gb < 0; gi++) {
        Group& g = *groups[gi];
        if (g.max_bw != cc.max_bw) continue;
        biggest = std::max(biggest, g.cap);
        for (int i = 0; i < g.cap; i++)
            if (i >= (int)g.chans.size() || g.chans[i] < 0) { gb = (int)gi; slot = i; break; }
    }
    if (gb < 0) {
        // no parked slot with this chain: a new group, with room for the next movers (capacity doubles per overflow group)
        const int room = std::min(next_pow2(nch), std::max(32, 2 * biggest));
        CSDR_TRY(create_group(cc.max_bw, std::vector<int>(), room, &gb));
        slot = 0;
    }
    Group& g = *groups[gb];
    if (slot >= (int)g.chans.size()) g.chans.resize(slot + 1, -1);
    g.chans[slot] = c;
    const int used = (int)g.chans.size();
    g.fir.set_nch(used);
    g.post.set_nch(used);
    if (g.rs) { g.rs->set_rows(used); CSDR_TRY(g.rs->reset_row(slot)); }
    CSDR_TRY(g.dec.reset_channel(slot));
    g.fir.release(slot);
    g.post.free_channel(slot);
    g.h_chan_map[slot] = c;
    CSDR_TRY(g.stage.upload(g.d_chan_map, g.h_chan_map.data(), g.cap * sizeof(int), g.st_post));
    cc.group = gb;
    cc.local = slot;
    CSDR_TRY(apply_channel(this, c, true));
    return CUTESDR_OK;
}

int cutesdr_bank::rebuild()
{
    CSDR_CK(cudaSetDevice(device));
    CSDR_CK(cudaStreamSynchronize(st));
    groups.clear();
    std::map<double, std::vector<int>> by_bw;
    for (int c = 0; c < nch; c++) {
        if (!ch[c].configured) { set_error("channel %d has no demodulator (call set_demod first)", c); return CUTESDR_E_STATE; }
        by_bw[ch[c].max_bw].push_back(c);
    }
    // DSP block length. The reference takes m_InBufLimit = 10 ms of input rounded down to a multiple of 256
    // (dsp/demodulator.cpp:145-146) although its decimator asks for a multiple of 2^stages
    // (dsp/downconvert.cpp:182-183): at e.g. exactly 100e6 sps (999 936 samples, 11 stages) its deeper stages then get
    // odd lengths and drop / re-read a sample at every block edge. The bank rounds the same 10 ms down to a multiple
    // of 2^stages of its deepest ladder instead, so ANY input rate is accepted and the decimated stream is the clean
    // one; whenever the reference's own length already is such a multiple the two are identical.
    int newL = -1, max_stages = 0;
    for (auto& kv : by_bw) {
        std::vector<int> lens;
        double orate = plan_stages(in_rate, kv.first, lens);
        int lim = block_limit(in_rate, orate);
        max_stages = std::max(max_stages, (int)lens.size());
        if (newL < 0) newL = lim;
        else if (lim != newL) { set_error("channel groups disagree on the DSP block length (%d vs %d)", lim, newL); return CUTESDR_E_ARG; }
    }
    if (newL > 0 && max_stages > 8) newL &= ~((1 << max_stages) - 1);
    if (newL <= 0) { set_error("input rate %g too low: DSP block length is %d", in_rate, newL); return CUTESDR_E_ARG; }
    if (newL != L) {
        L = newL;
        cudaFree(d_x);
        d_x = nullptr;
        if (h_stage) cudaFreeHost(h_stage);
        h_stage = nullptr;
        CSDR_CK(cudaMalloc(&d_x, (size_t)L * sizeof(float2)));
        for (int k = 0; k < kAsyncSlots; k++) { cudaFree(d_xs[k]); d_xs[k] = nullptr; }
        for (int k = 0; k < 2; k++) if (!d_halo[k]) CSDR_CK(cudaMalloc(&d_halo[k], (size_t)kHaloMax * sizeof(float2)));
        CSDR_CK(cudaHostAlloc(&h_stage, (size_t)L * sizeof(float2), cudaHostAllocDefault));
        h_fill = 0;
    }
    // a rebuild re-creates every DSP object: the stream restarts from zero state
    for (int k = 0; k < 2; k++) CSDR_CK(cudaMemsetAsync(d_halo[k], 0, (size_t)kHaloMax * sizeof(float2), st));
    halo_cur = 0;
    stream_pos = 0;
    block_index = 0;
    for (auto& kv : by_bw) {
        std::vector<int> sorted = kv.second;
        // inside a group channels are sorted by mode so warps of the lane-per-channel kernels do not diverge
        std::stable_sort(sorted.begin(), sorted.end(), [&](int a, int c2) { return ch[a].mode < ch[c2].mode; });
        CSDR_TRY(create_group(kv.first, sorted, 0, nullptr));
    }
    for (int c = 0; c < nch; c++) CSDR_TRY(apply_channel(this, c, true));
    if (nb_on || nb) {
        nb.reset(new Blanker());
        CSDR_TRY(nb->init(L, st, &lc));
        CSDR_TRY(nb->setup(nb_on, nb_thresh, nb_width, in_rate));
    }
    blk_nout.assign(nch, 0);
    layout_dirty = false;
    return CUTESDR_OK;
}

// One DSP block. d_block points at the block's first sample (kHaloMax history in front).
// audio_off[group] = samples already written for that group's channels in d_audio_out.
int cutesdr_bank::run_block(const void* d_block, int fmt, float* d_audio_out, int audio_stride, const int* audio_off, int* n_out_max)
{
    int nmax = 0;
    std::fill(blk_nout.begin(), blk_nout.end(), 0);
    for (size_t gi = 0; gi < groups.size(); gi++) {
        Group& g = *groups[gi];
        // The decimator may run ahead of the burst chain only as far as the 4096-sample ring allows:
        // the FIR window of a pending burst (2048 samples) must not be overwritten, which leaves 2048
        // samples of slack = `lag` further blocks of n_dec samples (one block is always in flight).
        const int lag = std::max(1, std::min(4, (kDecRing - kFirFft) / std::max(1, g.dec.out_per_block()) - 1));
        for (size_t k = 0; k < g.pending.size();) {
            if (g.pending[k].block <= block_index - lag) {
                CSDR_TRY(g.dec.wait_before_output(g.pending[k].ev));
                g.pending.erase(g.pending.begin() + k);
            } else k++;
        }
        CSDR_TRY(g.dec.run_block(d_block, d_halo[halo_cur], d_halo[halo_cur ^ 1], -1, fmt));
        const long long total = g.dec.total_out();
        const int nbursts = (int)(total / kBurst - g.bursts_done);
        g.last_fir_n = 0;
        if (nbursts <= 0) {
            if (!tap_spectra.empty()) CSDR_TRY(feed_tap_spectra(g, (int)gi, 0, nullptr, 0, 0));
            continue;
        }
        const int n = nbursts * kBurst;
        if (n > kMaxBurstSamples) { set_error("more than %d FIR bursts in one DSP block", kMaxBurstSamples / kBurst); return CUTESDR_E_STATE; }
        CSDR_CK(cudaStreamWaitEvent(g.st_post, g.dec.done_event(), 0));
        if (d2h_pending) CSDR_CK(cudaStreamWaitEvent(g.st_post, ev_d2h, 0));
        CSDR_TRY(g.fir.run(g.dec.ring(), g.bursts_done, nbursts, g.post.y_in(), g.post.y_stride()));
        g.bursts_done += nbursts;
        g.last_fir_n = n;
        const int off = audio_off ? audio_off[gi] : 0;
        int produced = n;
        if (g.rs) {
            // demod audio goes into the resampler's input rows, the resampler writes the user rows
            CSDR_TRY(g.post.run(n, g.rs->in_ptr(), g.rs->in_stride(), 0, g.d_local_map));
            const double rate = g.dec.out_rate() / audio_rate;     // interface/soundout.cpp:204
            if (d_audio_out && (stereo ? 2 : 1) * (off + g.rs->max_out(n, rate)) > audio_stride) {
                set_error("audio_stride %d too small for %d resampled samples at offset %d", audio_stride, g.rs->max_out(n, rate), off);
                return CUTESDR_E_ARG;
            }
            CSDR_TRY(g.rs->run(n, rate, d_audio_out, audio_stride, off, g.d_chan_map, &produced));
        } else {
            if (d_audio_out && (stereo ? 2 : 1) * (off + n) > audio_stride) {
                set_error("audio_stride %d too small for %d samples at offset %d", audio_stride, n, off);
                return CUTESDR_E_ARG;
            }
            CSDR_TRY(g.post.run(n, d_audio_out, audio_stride, off, g.d_chan_map));
        }
        if (!tap_spectra.empty()) CSDR_TRY(feed_tap_spectra(g, (int)gi, n, d_audio_out, audio_stride, off));
        cudaEvent_t done = g.ev_post[g.post_launches++ & 7];
        CSDR_CK(cudaEventRecord(done, g.st_post));
        g.pending.push_back({done, block_index});
        for (int c : g.chans) if (c >= 0) blk_nout[c] = produced;
        nmax = std::max(nmax, produced);
    }
    halo_cur ^= 1;      // kernel 1 saved this block's tail into the other halo buffer
    last_block = d_block;
    last_fmt = fmt;
    stream_pos += L;
    block_index++;
    if (n_out_max) *n_out_max = nmax;
    return CUTESDR_OK;
}

int cutesdr_bank::join()
{
    for (auto& g : groups) {
        CSDR_TRY(g->dec.join_main());
        for (auto& p : g->pending) CSDR_CK(cudaStreamWaitEvent(st, p.ev, 0));
        g->pending.clear();
    }
    return CUTESDR_OK;
}

int cutesdr_bank::sync_all()
{
    CSDR_TRY(join());
    CSDR_CK(cudaStreamSynchronize(st));
    if (st_h2d) CSDR_CK(cudaStreamSynchronize(st_h2d));
    if (st_d2h) CSDR_CK(cudaStreamSynchronize(st_d2h));
    d2h_pending = false;
    return CUTESDR_OK;
}

// CTestBench::DisplayData, frequency-domain branch (gui/testbench.cpp:594-611), for n samples at src on stream st:
// fill m_FftInBuf; every full frame bumps the skip counter and, when it reaches m_DisplaySkipValue, goes through
// PutInDisplayFFT of the attached CFft. All of it is queued in stream order; nothing synchronises.
static int tap_spectrum_append(cutesdr_bank* b, TapSpectrum& t, const void* src, int n, int kind, cudaStream_t st)
{
    const size_t esz = kind ? sizeof(float) : sizeof(float2);
    const unsigned char* p = reinterpret_cast<const unsigned char*>(src);
    while (n > 0) {
        const int take = std::min(n, kTestFftSize - t.pos);
        k_tap_append<<<(take + 255) / 256, 256, 0, st>>>(t.d_frame + t.pos, p, take, kind);
        b->lc.n++;
        CSDR_CK(cudaGetLastError());
        t.pos += take;
        p += (size_t)take * esz;
        n -= take;
        if (t.pos >= kTestFftSize) {
            t.pos = 0;
            if (++t.skip_counter >= t.skip_value) {
                t.skip_counter = 0;
                int total = 0;
                CSDR_TRY(cutesdr_fft_put_device_async(t.fft, kTestFftSize, t.d_frame, (void*)st, &total));
                t.frames++;
            }
        }
    }
    return CUTESDR_OK;
}

// Feed the attached test-bench spectra of group g for the block just queued: PROFILE_1 = this block's decimated
// samples (dsp/demodulator.cpp:175), PROFILE_2/3/4 = the burst's FIR output, AGC output and audio (:180,187,208).
// Runs on the group's burst stream, after the burst chain.
int cutesdr_bank::feed_tap_spectra(Group& g, int gi, int n_burst, float* d_audio_out, int audio_stride, int audio_off)
{
    for (auto& tp : tap_spectra) {
        TapSpectrum& t = *tp;
        const ChanCfg& cc = ch[t.ch];
        if (cc.group != gi) continue;
        const int i = cc.local;
        if (t.profile == 1) {
            const int n_dec = g.dec.out_per_block();
            CSDR_CK(cudaStreamWaitEvent(g.st_post, g.dec.done_event(), 0));
            const long long first = g.dec.total_out() - n_dec;
            for (int k = 0; k < n_dec;) {
                const int pos = (int)((first + k) & (kDecRing - 1));
                const int run = std::min(n_dec - k, kDecRing - pos);
                CSDR_TRY(tap_spectrum_append(this, t, g.dec.ring() + (size_t)i * kDecRing + pos, run, 0, g.st_post));
                k += run;
            }
            // the ring must not be overwritten before these reads: same ordering as a pending burst chain
            cudaEvent_t& e = t.ev[t.n_ev++ & 7];
            if (!e) CSDR_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            CSDR_CK(cudaEventRecord(e, g.st_post));
            g.pending.push_back({e, block_index});
            continue;
        }
        if (n_burst <= 0) continue;
        if (t.profile == 2)
            CSDR_TRY(tap_spectrum_append(this, t, g.post.y_in() + (size_t)i * g.post.y_stride() - n_burst, n_burst, 0, g.st_post));
        else if (t.profile == 3)
            CSDR_TRY(tap_spectrum_append(this, t, g.post.tap3() + (size_t)i * g.post.tap3_stride(), n_burst, 0, g.st_post));
        else if (t.profile == 4) {
            // the demodulator output at m_OutputRate, i.e. in front of the bank resampler; stereo frames are complex
            const float* src = nullptr;
            if (g.rs) src = g.rs->in_ptr() + (size_t)i * g.rs->in_stride();
            else if (d_audio_out) src = d_audio_out + (size_t)t.ch * audio_stride + (size_t)(stereo ? 2 : 1) * audio_off;
            if (src) CSDR_TRY(tap_spectrum_append(this, t, src, n_burst, stereo ? 0 : 1, g.st_post));
        }
    }
    return CUTESDR_OK;
}

// slow path: copy the enabled test-bench taps of the block just processed to the host
int cutesdr_bank::collect_taps()
{
    bool any = false;
    for (auto& g : groups) any |= g->any_tap;
    if (!any) return CUTESDR_OK;
    CSDR_TRY(sync_all());
    for (auto& gp : groups) {
        Group& g = *gp;
        if (!g.any_tap) continue;
        const int n_dec = g.dec.out_per_block();
        for (int i = 0; i < (int)g.chans.size(); i++) {
            if (g.chans[i] < 0) continue;
            ChanCfg& cc = ch[g.chans[i]];
            if (!cc.tap_mask) continue;
            if (cc.tap_mask & 2u) {   // PROFILE_1: decimated samples of this block
                std::vector<float2> tmp(n_dec);
                long long first = g.dec.total_out() - n_dec;
                for (int k = 0; k < n_dec;) {
                    int pos = (int)((first + k) & (kDecRing - 1));
                    int run = std::min(n_dec - k, kDecRing - pos);
                    CSDR_CK(cudaMemcpy(tmp.data() + k, g.dec.ring() + (size_t)i * kDecRing + pos, run * sizeof(float2), cudaMemcpyDeviceToHost));
                    k += run;
                }
                const float* f = reinterpret_cast<const float*>(tmp.data());
                cc.tap[1].insert(cc.tap[1].end(), f, f + 2 * n_dec);
            }
            const int n = g.last_fir_n;
            if (n > 0 && (cc.tap_mask & 4u)) {   // PROFILE_2: FIR output row (the post stage has already shifted its
                                                 // delay history, so the burst now ends at y_in()-... : read it from there)
                std::vector<float2> tmp(n);
                CSDR_CK(cudaMemcpy(tmp.data(), g.post.y_in() + (size_t)i * g.post.y_stride() - n, n * sizeof(float2), cudaMemcpyDeviceToHost));
                const float* f = reinterpret_cast<const float*>(tmp.data());
                cc.tap[2].insert(cc.tap[2].end(), f, f + 2 * n);
            }
            if (n > 0 && (cc.tap_mask & 8u)) {   // PROFILE_3: post-AGC row
                std::vector<float2> tmp(n);
                CSDR_CK(cudaMemcpy(tmp.data(), g.post.tap3() + (size_t)i * g.post.tap3_stride(), n * sizeof(float2), cudaMemcpyDeviceToHost));
                const float* f = reinterpret_cast<const float*>(tmp.data());
                cc.tap[3].insert(cc.tap[3].end(), f, f + 2 * n);
            }
        }
    }
    return CUTESDR_OK;
}

extern "C" {

const char* cutesdr_last_error(void) { return g_err.c_str(); }
const char* cutesdr_version(void) { return "cutesdr_b200 0.1 (sm_100a)"; }

int cutesdr_device_count(int* n)
{
    int k = 0;
    cudaError_t e = cudaGetDeviceCount(&k);
    if (e != cudaSuccess) { set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e)); if (n) *n = 0; return CUTESDR_E_CUDA; }
    if (n) *n = k;
    return CUTESDR_OK;
}

int cutesdr_host_alloc(void** p, size_t bytes)
{
    if (!p || bytes == 0) { set_error("host_alloc: bad arguments"); return CUTESDR_E_ARG; }
    *p = nullptr;
    CSDR_CK(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return CUTESDR_OK;
}

void cutesdr_host_free(void* p)
{
    if (p) cudaFreeHost(p);
}

int cutesdr_device_memory(int device, long long* free_bytes, long long* total_bytes)
{
    CSDR_CK(cudaSetDevice(device));
    size_t f = 0, t = 0;
    CSDR_CK(cudaMemGetInfo(&f, &t));
    if (free_bytes) *free_bytes = (long long)f;
    if (total_bytes) *total_bytes = (long long)t;
    return CUTESDR_OK;
}

int cutesdr_bank_create(cutesdr_bank** out, int n_channels, double in_rate, int device)
{
    if (!out || n_channels <= 0 || !(in_rate > 0)) { set_error("bank_create: bad arguments"); return CUTESDR_E_ARG; }
    *out = nullptr;
    CSDR_CK(cudaSetDevice(device));
    std::unique_ptr<cutesdr_bank> b(new cutesdr_bank());
    b->nch = n_channels;
    b->in_rate = in_rate;
    b->device = device;
    b->ch.resize(n_channels);
    CSDR_CK(cudaStreamCreateWithFlags(&b->st, cudaStreamNonBlocking));
    *out = b.release();
    return CUTESDR_OK;
}

void cutesdr_bank_destroy(cutesdr_bank* b)
{
    if (!b) return;
    cudaSetDevice(b->device);
    delete b;
}

int cutesdr_bank_set_demod(cutesdr_bank* b, int c, int mode, const cutesdr_demod_info* info)
{
    if (!b || !info || c < 0 || c >= b->nch || mode < 0 || mode > CUTESDR_DEMOD_CWL) { set_error("set_demod: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    ChanCfg& cc = b->ch[c];
    cc.info = *info;                                   // dsp/demodulator.cpp:110
    bool new_demod = false, chain_changed = false;
    if (cc.mode != mode) {                             // :111-142
        cc.mode = mode;
        new_demod = true;
        if (mode == CUTESDR_DEMOD_LSB || mode == CUTESDR_DEMOD_CWL) cc.max_bw = -cc.info.LowCutmin;
        else cc.max_bw = cc.info.HiCutmax;
        if (cc.dc_max_bw != cc.max_bw) {
            // CDownConvert::SetDataRate rebuilds the chain and ends with SetFrequency(m_NcoFreq),
            // which adds the current CW offset once more (dsp/downconvert.cpp:98-107,168)
            cc.dc_max_bw = cc.max_bw;
            cc.dc_nco_freq = cc.dc_nco_freq + cc.dc_cw;
            chain_changed = true;
        }
    }
    cc.demod_cw = cc.info.Offset;                      // :143-144
    cc.dc_cw = cc.demod_cw;
    const bool first = !cc.configured;
    cc.configured = true;
    if (first) b->layout_dirty = true;
    if (chain_changed && !b->layout_dirty && cc.group >= 0) {
        cc.demod_cw = cc.info.Offset;
        CSDR_TRY(b->move_channel(c));                  // restarts this channel only
        return CUTESDR_OK;
    }
    if (chain_changed) b->layout_dirty = true;
    if (!b->layout_dirty) CSDR_TRY(apply_channel(b, c, new_demod));
    return CUTESDR_OK;
}

int cutesdr_bank_set_demod_freq(cutesdr_bank* b, int c, double freq)
{
    if (!b || c < 0 || c >= b->nch) { set_error("set_demod_freq: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    ChanCfg& cc = b->ch[c];
    cc.dc_cw = cc.demod_cw;                            // dsp/demodulator.h:68
    cc.dc_nco_freq = freq + cc.dc_cw;                  // dsp/downconvert.cpp:100-102
    if (!b->layout_dirty && cc.group >= 0) b->groups[cc.group]->dec.set_frequency(cc.local, cc.dc_nco_freq);
    return CUTESDR_OK;
}

int cutesdr_bank_get_output_rate(cutesdr_bank* b, int c, double* rate)
{
    if (!b || !rate || c < 0 || c >= b->nch) { set_error("get_output_rate: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    std::vector<int> lens;
    *rate = plan_stages(b->in_rate, b->ch[c].dc_max_bw, lens);
    return CUTESDR_OK;
}

int cutesdr_bank_block_length(cutesdr_bank* b, int* n)
{
    if (!b || !n) { set_error("block_length: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    *n = b->L;
    return CUTESDR_OK;
}

int cutesdr_bank_get_smeter(cutesdr_bank* b, int c, double* peak, double* ave)
{
    if (!b || c < 0 || c >= b->nch) { set_error("get_smeter: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    ChanCfg& cc = b->ch[c];
    return b->groups[cc.group]->post.read_smeter(cc.local, peak, ave);
}

int cutesdr_bank_set_noiseproc(cutesdr_bank* b, int on, double threshold, double width_us)
{
    if (!b) { set_error("set_noiseproc: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    b->nb_on = on != 0;
    b->nb_thresh = threshold;
    b->nb_width = width_us;
    if (!b->layout_dirty) {
        if (!b->nb) { b->nb.reset(new Blanker()); CSDR_TRY(b->nb->init(b->L, b->st, &b->lc)); }
        CSDR_TRY(b->nb->setup(b->nb_on, threshold, width_us, b->in_rate));
    }
    return CUTESDR_OK;
}

int cutesdr_bank_set_stereo(cutesdr_bank* b, int stereo)
{
    if (!b) { set_error("set_stereo: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    if ((stereo != 0) != b->stereo) { b->stereo = stereo != 0; b->layout_dirty = true; }
    return CUTESDR_OK;
}

int cutesdr_bank_set_audio_rate(cutesdr_bank* b, double audio_rate)
{
    if (!b) { set_error("set_audio_rate: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    if (audio_rate != b->audio_rate) { b->audio_rate = audio_rate > 0 ? audio_rate : 0.0; b->layout_dirty = true; }
    return CUTESDR_OK;
}

static int ensure_audio(cutesdr_bank* b, int stride)
{
    if (stride <= b->audio_cap) return CUTESDR_OK;
    cudaFree(b->d_audio);
    b->d_audio = nullptr;
    b->audio_cap = 0;
    CSDR_CK(cudaMalloc(&b->d_audio, (size_t)b->nch * stride * sizeof(float)));
    b->audio_cap = stride;
    return CUTESDR_OK;
}

// Wideband pre-processing shared by all channels. Returns (in *blk) the device block kernel 1 reads:
// the caller's own device buffer when nothing has to touch it, else the bank's staging buffer
// (host input, the noise blanker's output, or the first block of a stream whose first samples get the
// oscillator's start-up amplitude).
static int stage_block(cutesdr_bank* b, const void* src, int fmt, cudaMemcpyKind kind, const void** blk, int* blk_fmt)
{
    const bool nb_on = b->nb && b->nb->on();
    const bool startup = b->stream_pos < kNcoStartup;
    const size_t bytes = (size_t)b->L * sample_bytes(fmt);
    *blk_fmt = 0;
    if (!nb_on && !startup) {
        // fast paths: kernel 1 reads the block where it lies, in whatever format it has
        *blk_fmt = fmt;
        if (kind == cudaMemcpyDeviceToDevice) { *blk = src; return CUTESDR_OK; }
        CSDR_CK(cudaMemcpyAsync(b->d_x, src, bytes, kind, b->st));
        *blk = b->d_x;
        return CUTESDR_OK;
    }
    // slow paths need complex64 in a scratch buffer first
    float2* f32 = nullptr;
    const float2* in = nullptr;
    if (kind == cudaMemcpyDeviceToDevice && fmt == 0) in = reinterpret_cast<const float2*>(src);
    else {
        CSDR_CK(cudaMallocAsync(&f32, (size_t)b->L * sizeof(float2), b->st));
        if (fmt == 0) CSDR_CK(cudaMemcpyAsync(f32, src, bytes, kind, b->st));
        else {
            const void* raw = src;
            void* tmp = nullptr;
            if (kind != cudaMemcpyDeviceToDevice) {
                CSDR_CK(cudaMallocAsync(&tmp, bytes, b->st));
                CSDR_CK(cudaMemcpyAsync(tmp, src, bytes, kind, b->st));
                raw = tmp;
            }
            CSDR_TRY(unpack_samples(raw, fmt, f32, b->L, b->st, &b->lc));
            if (tmp) CSDR_CK(cudaFreeAsync(tmp, b->st));
        }
        in = f32;
    }
    int rc = CUTESDR_OK;
    if (nb_on) rc = b->nb->run(in, b->d_x, b->L);
    else CSDR_CK(cudaMemcpyAsync(b->d_x, in, (size_t)b->L * sizeof(float2), cudaMemcpyDeviceToDevice, b->st));
    if (f32) CSDR_CK(cudaFreeAsync(f32, b->st));
    if (rc < 0) return rc;
    if (startup) CSDR_TRY(apply_nco_startup_gain(b->d_x, b->stream_pos, b->L, b->st, &b->lc));
    *blk = b->d_x;
    return CUTESDR_OK;
}

static int bank_process_host(cutesdr_bank* b, int n_in, const void* data, int fmt, float* audio, int audio_stride, int* n_out);

int cutesdr_bank_process_raw(cutesdr_bank* b, int n_in, const void* data, int fmt, float* audio, int audio_stride, int* n_out)
{
    if (fmt < 0 || fmt > 2) { set_error("bank_process_raw: unknown sample format %d", fmt); return CUTESDR_E_ARG; }
    return bank_process_host(b, n_in, data, fmt, audio, audio_stride, n_out);
}

int cutesdr_bank_process(cutesdr_bank* b, int n_in, const float* iq, float* audio, int audio_stride, int* n_out)
{
    return bank_process_host(b, n_in, iq, 0, audio, audio_stride, n_out);
}

// CUdpThread::OnreadyRead (interface/netiobase.cpp:464-534) + CIQDataThread::run (:571-600): datagrams of 1028 bytes
// (256 int16 I/Q pairs) or 1444 bytes (240 packed int24 pairs) after a 4-byte header whose bytes 2-3 are a little-endian
// sequence number. The sequence bookkeeping is the reference's, statement for statement (16-bit wrap skips 0, a 0 from
// the radio restarts the count, the gap is added as a signed 16-bit difference); payloads go on in their wire format,
// so the unpack still happens inside kernel 1's tile load.
int cutesdr_bank_process_packets(cutesdr_bank* b, int n_packets, const void* packets, int packet_bytes, float* audio,
                                 int audio_stride, int* n_out)
{
    if (!b || n_packets < 0 || (n_packets > 0 && !packets)) { set_error("bank_process_packets: bad arguments"); return CUTESDR_E_ARG; }
    int fmt, payload;
    if (packet_bytes == CUTESDR_PKT_LENGTH_16) { fmt = CUTESDR_FMT_CS16; payload = CUTESDR_PKT_LENGTH_16 - 4; }
    else if (packet_bytes == CUTESDR_PKT_LENGTH_24) { fmt = CUTESDR_FMT_CS24; payload = CUTESDR_PKT_LENGTH_24 - 4; }
    else {
        // the reference silently ignores datagrams of any other size (:480,:507); a whole call of them is a caller error
        set_error("bank_process_packets: packet length %d is neither %d (16-bit) nor %d (24-bit)", packet_bytes,
                  CUTESDR_PKT_LENGTH_16, CUTESDR_PKT_LENGTH_24);
        return CUTESDR_E_ARG;
    }
    {
        std::lock_guard<std::mutex> lk(b->mu);
        b->pkt_payload.resize((size_t)n_packets * payload);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(packets);
        for (int k = 0; k < n_packets; k++, src += packet_bytes) {
            const unsigned short seq = (unsigned short)(src[2] | (src[3] << 8));
            if (0 == seq) b->pkt_last_seq = 0;                         // first packet after the radio started
            if (seq != b->pkt_last_seq) {
                b->missed_packets += (short)seq - (short)b->pkt_last_seq;
                b->pkt_last_seq = seq;
            }
            b->pkt_last_seq++;
            if (0 == b->pkt_last_seq) b->pkt_last_seq = 1;
            memcpy(b->pkt_payload.data() + (size_t)k * payload, src + 4, payload);
        }
    }
    const int samples = n_packets * (fmt == CUTESDR_FMT_CS16 ? payload / 4 : payload / 6);
    return bank_process_host(b, samples, b->pkt_payload.data(), fmt, audio, audio_stride, n_out);
}

int cutesdr_bank_missed_packets(cutesdr_bank* b, long long* missed, int reset)
{
    if (!b) { set_error("missed_packets: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    if (missed) *missed = b->missed_packets;
    if (reset) { b->missed_packets = 0; b->pkt_last_seq = 0; }
    return CUTESDR_OK;
}

static int bank_process_host(cutesdr_bank* b, int n_in, const void* iq, int fmt, float* audio, int audio_stride, int* n_out)
{
    if (!b || n_in < 0 || (n_in > 0 && !iq) || (audio && audio_stride <= 0)) { set_error("bank_process: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    if (audio) CSDR_TRY(ensure_audio(b, audio_stride));
    std::vector<int> goff(b->groups.size(), 0);
    std::vector<int> nout(b->nch, 0);
    const int sb = sample_bytes(fmt);
    if (b->h_fill > 0 && fmt != b->h_fmt) { set_error("bank_process: sample format changed inside a partially filled block"); return CUTESDR_E_STATE; }
    b->h_fmt = fmt;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(iq);
    int pos = 0, nmax = 0;
    while (pos < n_in) {
        const void* blk = nullptr;
        if (b->h_fill == 0 && n_in - pos >= b->L) {      // whole block available: no staging copy
            blk = src + (size_t)pos * sb;
            pos += b->L;
        } else {
            int take = std::min(n_in - pos, b->L - b->h_fill);
            memcpy(reinterpret_cast<unsigned char*>(b->h_stage) + (size_t)b->h_fill * sb, src + (size_t)pos * sb, (size_t)take * sb);
            b->h_fill += take;
            pos += take;
            if (b->h_fill < b->L) break;
            blk = b->h_stage;
            b->h_fill = 0;
        }
        const void* dblk = nullptr;
        int dfmt = 0;
        CSDR_TRY(stage_block(b, blk, fmt, cudaMemcpyHostToDevice, &dblk, &dfmt));
        int m = 0;
        CSDR_TRY(b->run_block(dblk, dfmt, audio ? b->d_audio : nullptr, b->audio_cap, goff.data(), &m));
        for (size_t gi = 0; gi < b->groups.size(); gi++) {
            Group& g = *b->groups[gi];
            int first = -1;
            for (int c : g.chans) if (c >= 0) { first = c; break; }
            if (first < 0) continue;
            int produced = b->blk_nout[first];
            if (produced > 0 && audio) {
                // PROFILE_4 tap = the audio rows themselves
                for (int c : g.chans) if (c >= 0 && (b->ch[c].tap_mask & 16u)) {
                    CSDR_TRY(b->sync_all());
                    const int w = b->stereo ? 2 : 1;
                    std::vector<float> tmp((size_t)w * produced);
                    CSDR_CK(cudaMemcpyAsync(tmp.data(), b->d_audio + (size_t)c * b->audio_cap + (size_t)w * goff[gi], (size_t)w * produced * sizeof(float), cudaMemcpyDeviceToHost, b->st));
                    CSDR_CK(cudaStreamSynchronize(b->st));
                    b->ch[c].tap[4].insert(b->ch[c].tap[4].end(), tmp.begin(), tmp.end());
                }
            }
            goff[gi] += produced;
            for (int c : g.chans) if (c >= 0) nout[c] += produced;
            nmax = std::max(nmax, goff[gi]);
        }
        CSDR_TRY(b->collect_taps());
        // the staging buffer (or the caller's memory) must not change until the copy is done
        CSDR_CK(cudaStreamSynchronize(b->st));
    }
    CSDR_TRY(b->join());
    if (audio && nmax > 0) {
        CSDR_CK(cudaMemcpy2DAsync(audio, (size_t)audio_stride * sizeof(float), b->d_audio, (size_t)b->audio_cap * sizeof(float),
                                  (size_t)nmax * (b->stereo ? 2 : 1) * sizeof(float), b->nch, cudaMemcpyDeviceToHost, b->st));
    }
    CSDR_CK(cudaStreamSynchronize(b->st));
    if (n_out) memcpy(n_out, nout.data(), b->nch * sizeof(int));
    return nmax;
}

static int bank_process_async(cutesdr_bank* b, int n_in, const void* iq, int fmt, float* audio, int audio_stride, int* n_out,
                              bool from_device = false, cudaStream_t src_stream = 0, cutesdr_mgpu* mg = nullptr);

int cutesdr_bank_process_async(cutesdr_bank* b, int n_in, const float* iq, float* audio, int audio_stride, int* n_out)
{
    return bank_process_async(b, n_in, iq, 0, audio, audio_stride, n_out);
}

int cutesdr_bank_process_async_raw(cutesdr_bank* b, int n_in, const void* data, int fmt, float* audio, int audio_stride, int* n_out)
{
    if (fmt < 0 || fmt > 2) { set_error("bank_process_async_raw: unknown sample format %d", fmt); return CUTESDR_E_ARG; }
    return bank_process_async(b, n_in, data, fmt, audio, audio_stride, n_out);
}

int cutesdr_bank_process_async_device(cutesdr_bank* b, int n_in, const void* d_iq, void* src_stream, float* audio, int audio_stride, int* n_out)
{
    return bank_process_async(b, n_in, d_iq, 0, audio, audio_stride, n_out, true, reinterpret_cast<cudaStream_t>(src_stream));
}

int cutesdr_bank_process_async_bcast(cutesdr_bank* b, cutesdr_mgpu* m, int n_in, const void* iq_rank0, int fmt, float* audio,
                                     int audio_stride, int* n_out)
{
    if (!m || fmt < 0 || fmt > 2) { set_error("bank_process_async_bcast: bad arguments"); return CUTESDR_E_ARG; }
    if (b && b->device != m->device) { set_error("bank_process_async_bcast: bank and communicator live on different devices"); return CUTESDR_E_ARG; }
    return bank_process_async(b, n_in, iq_rank0, fmt, audio, audio_stride, n_out, false, 0, m);
}

static int bank_process_async(cutesdr_bank* b, int n_in, const void* iq, int fmt, float* audio, int audio_stride, int* n_out,
                              bool from_device, cudaStream_t src_stream, cutesdr_mgpu* mg)
{
    const bool need_iq = !mg || mg->rank == 0;
    if (!b || (need_iq && !iq) || (audio && audio_stride <= 0)) { set_error("bank_process_async: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    if (n_in != b->L || b->h_fill != 0) {
        set_error("bank_process_async: n_in %d must equal the block length %d (and no partial block may be pending)", n_in, b->L);
        return CUTESDR_E_ARG;
    }
    if (audio) CSDR_TRY(ensure_audio(b, audio_stride));
    if (!b->st_h2d) {
        CSDR_CK(cudaStreamCreateWithFlags(&b->st_h2d, cudaStreamNonBlocking));
        CSDR_CK(cudaStreamCreateWithFlags(&b->st_d2h, cudaStreamNonBlocking));
        for (int k = 0; k < kAsyncSlots; k++) {
            CSDR_CK(cudaEventCreateWithFlags(&b->ev_h2d[k], cudaEventDisableTiming));
            CSDR_CK(cudaEventCreateWithFlags(&b->ev_free[k], cudaEventDisableTiming));
            CSDR_CK(cudaEventCreateWithFlags(&b->ev_host[k], cudaEventDisableTiming));
        }
        CSDR_CK(cudaEventCreateWithFlags(&b->ev_d2h, cudaEventDisableTiming));
    }
    const int slot = (int)(b->async_blocks % kAsyncSlots);
    if (!b->d_xs[slot]) CSDR_CK(cudaMalloc(&b->d_xs[slot], (size_t)b->L * sizeof(float2)));
    // The header's promise ("iq must stay unchanged until the second following call") is enforced here: the host waits
    // for the copy that read the buffer handed in two calls ago before this call returns, so a producer that runs ahead
    // of the GPU can never overwrite samples the DMA has not read yet.
    if (b->async_blocks >= 2 && !from_device && need_iq) CSDR_CK(cudaEventSynchronize(b->ev_host[(b->async_blocks - 2) % kAsyncSlots]));
    // the transfer into this slot starts as soon as the block that used it kAsyncSlots calls ago has been consumed
    if (b->async_blocks >= kAsyncSlots) {
        CSDR_CK(cudaStreamWaitEvent(b->st_h2d, b->ev_free[slot], 0));
        if (mg) CSDR_CK(cudaStreamWaitEvent(mg->st_comm, b->ev_free[slot], 0));
    }
    if (mg) {
        // multi-GPU: rank 0's pinned host block goes H2D in chunks, every chunk is broadcast over NVLink as soon as it
        // has landed, straight into this slot on every rank; ev_h2d[slot] fires when the whole block is here
        std::lock_guard<std::mutex> lk2(mg->mu);
        CSDR_TRY(mgpu_bcast_block(mg, iq, b->d_xs[slot], (size_t)b->L * sample_bytes(fmt), b->st_h2d, b->ev_h2d[slot]));
        CSDR_CK(cudaEventRecord(b->ev_host[slot], b->st_h2d));          // after the last host -> device chunk
    } else if (from_device) {
        // the block sits in the caller's device buffer and is read in stream order with respect to src_stream: the
        // copy into the slot waits for everything queued there so far (an NCCL broadcast, ...), and src_stream
        // waits for the copy, so the caller may queue the next write into the same buffer straight away
        CSDR_CK(cudaEventRecord(b->ev_h2d[slot], src_stream));
        CSDR_CK(cudaStreamWaitEvent(b->st_h2d, b->ev_h2d[slot], 0));
        CSDR_CK(cudaMemcpyAsync(b->d_xs[slot], iq, (size_t)b->L * sizeof(float2), cudaMemcpyDeviceToDevice, b->st_h2d));
        CSDR_CK(cudaEventRecord(b->ev_h2d[slot], b->st_h2d));
        CSDR_CK(cudaStreamWaitEvent(src_stream, b->ev_h2d[slot], 0));
    } else {
        CSDR_CK(cudaMemcpyAsync(b->d_xs[slot], iq, (size_t)b->L * sample_bytes(fmt), cudaMemcpyHostToDevice, b->st_h2d));
        CSDR_CK(cudaEventRecord(b->ev_h2d[slot], b->st_h2d));
        CSDR_CK(cudaEventRecord(b->ev_host[slot], b->st_h2d));
    }
    CSDR_CK(cudaStreamWaitEvent(b->st, b->ev_h2d[slot], 0));
    const void* dblk = nullptr;
    int dfmt = 0;
    CSDR_TRY(stage_block(b, b->d_xs[slot], fmt, cudaMemcpyDeviceToDevice, &dblk, &dfmt));
    int m = 0;
    std::vector<int> goff(b->groups.size(), 0);
    CSDR_TRY(b->run_block(dblk, dfmt, audio ? b->d_audio : nullptr, b->audio_cap, goff.data(), &m));
    CSDR_CK(cudaEventRecord(b->ev_free[slot], b->st));
    b->async_blocks++;
    if (n_out) for (int c = 0; c < b->nch; c++) n_out[c] = b->blk_nout[c];
    if (audio && m > 0) {
        // D2H of the finished audio rows on its own stream, after every group's burst chain
        for (auto& g : b->groups)
            for (auto& p : g->pending) if (p.block == b->block_index - 1) CSDR_CK(cudaStreamWaitEvent(b->st_d2h, p.ev, 0));
        CSDR_CK(cudaMemcpy2DAsync(audio, (size_t)audio_stride * sizeof(float), b->d_audio, (size_t)b->audio_cap * sizeof(float),
                                  (size_t)m * (b->stereo ? 2 : 1) * sizeof(float), b->nch, cudaMemcpyDeviceToHost, b->st_d2h));
        CSDR_CK(cudaEventRecord(b->ev_d2h, b->st_d2h));
        b->d2h_pending = true;
    }
    return m;
}

int cutesdr_bank_process_device(cutesdr_bank* b, const void* d_iq, int n_in, void* d_audio, int audio_stride, int* n_out_max)
{
    return cutesdr_bank_process_device_raw(b, d_iq, 0, n_in, d_audio, audio_stride, n_out_max);
}

int cutesdr_bank_process_device_raw(cutesdr_bank* b, const void* d_iq, int fmt, int n_in, void* d_audio, int audio_stride, int* n_out_max)
{
    if (!b || !d_iq || fmt < 0 || fmt > 2) { set_error("bank_process_device: bad arguments"); return CUTESDR_E_ARG; }
    // kernel 1 reads complex64 / int16 blocks with 16-byte vector loads
    if (fmt != 2 && (reinterpret_cast<uintptr_t>(d_iq) & 15u)) { set_error("bank_process_device: the device block must be 16-byte aligned"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    if (n_in != b->L) { set_error("bank_process_device: n_in %d must equal the block length %d", n_in, b->L); return CUTESDR_E_ARG; }
    const void* dblk = nullptr;
    int dfmt = 0;
    CSDR_TRY(stage_block(b, d_iq, fmt, cudaMemcpyDeviceToDevice, &dblk, &dfmt));
    int m = 0;
    CSDR_TRY(b->run_block(dblk, dfmt, reinterpret_cast<float*>(d_audio), audio_stride, nullptr, &m));
    CSDR_TRY(b->collect_taps());
    if (n_out_max) *n_out_max = m;
    return m;
}

int cutesdr_bank_synchronize(cutesdr_bank* b)
{
    if (!b) { set_error("synchronize: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    return b->sync_all();
}

int cutesdr_bank_join(cutesdr_bank* b)
{
    if (!b) { set_error("join: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    return b->join();
}

int cutesdr_bank_last_block(cutesdr_bank* b, const void** d_block, int* n)
{
    if (!b || !d_block) { set_error("last_block: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    if (b->last_fmt != 0) { set_error("last_block: the last block is still in its raw integer format"); return CUTESDR_E_STATE; }
    *d_block = b->last_block;
    if (n) *n = b->L;
    return b->last_block ? CUTESDR_OK : CUTESDR_E_STATE;
}

int cutesdr_bank_stream(cutesdr_bank* b, void** stream)
{
    if (!b || !stream) { set_error("bank_stream: bad arguments"); return CUTESDR_E_ARG; }
    *stream = (void*)b->st;
    return CUTESDR_OK;
}

int cutesdr_bank_launch_count(cutesdr_bank* b, long long* n)
{
    if (!b || !n) { set_error("launch_count: bad arguments"); return CUTESDR_E_ARG; }
    *n = b->lc.n;
    return CUTESDR_OK;
}

int cutesdr_bank_kernel_timing(cutesdr_bank* b, int enable)
{
    if (!b) { set_error("kernel_timing: bad handle"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    for (auto& g : b->groups) g->dec.enable_timing(enable != 0);
    return CUTESDR_OK;
}

int cutesdr_bank_kernel_time(cutesdr_bank* b, int which, double* ms_total, long long* launches)
{
    if (!b || which != 0) { set_error("kernel_time: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    double ms = 0.0;
    long long n = 0;
    for (auto& g : b->groups) {
        double m = 0.0;
        long long k = 0;
        CSDR_TRY(g->dec.read_timing(&m, &k));
        ms += m;
        n += k;
    }
    if (ms_total) *ms_total = ms;
    if (launches) *launches = n;
    return CUTESDR_OK;
}

int cutesdr_bank_kernel_model(cutesdr_bank* b, int which, int* on_tensor_cores, double* flops_per_block)
{
    if (!b || which < 0 || which > 1) { set_error("kernel_model: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    int all = 0;
    double fl = 0.0;
    bool any = false;
    for (auto& g : b->groups) {
        bool used = false;                  // groups whose slots are all parked run no kernels
        for (int u : g->chans) used |= (u >= 0);
        if (!used) continue;
        if (!any) { any = true; all = 1; }
        if (!g->dec.tensor_path() || (which == 1 && !g->dec.tensor_f16())) all = 0;
        fl += g->dec.tensor_flops_per_block();
    }
    if (on_tensor_cores) *on_tensor_cores = all;
    if (flops_per_block) *flops_per_block = fl;
    return CUTESDR_OK;
}

int cutesdr_bank_tap_enable(cutesdr_bank* b, int c, unsigned profile_mask)
{
    if (!b || c < 0 || c >= b->nch) { set_error("tap_enable: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    b->ch[c].tap_mask = profile_mask;
    for (int p = 1; p <= 4; p++) b->ch[c].tap[p].clear();
    for (auto& g : b->groups) {
        g->any_tap = false;
        for (int u : g->chans) if (u >= 0 && (b->ch[u].tap_mask & 0xEu)) g->any_tap = true;
    }
    return CUTESDR_OK;
}

int cutesdr_bank_tap_spectrum(cutesdr_bank* b, int c, int profile, cutesdr_fft* fft, int display_rate)
{
    if (!b || c < 0 || c >= b->nch || profile < 1 || profile > 4 || (fft && display_rate <= 0)) { set_error("tap_spectrum: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    CSDR_CK(cudaSetDevice(b->device));
    if (b->layout_dirty) CSDR_TRY(b->rebuild());
    // one spectrum per (channel, profile); a NULL fft detaches
    for (size_t k = 0; k < b->tap_spectra.size();) {
        if (b->tap_spectra[k]->ch == c && b->tap_spectra[k]->profile == profile) {
            CSDR_TRY(b->sync_all());
            b->tap_spectra.erase(b->tap_spectra.begin() + k);
        } else k++;
    }
    if (!fft) return CUTESDR_OK;
    std::unique_ptr<TapSpectrum> t(new TapSpectrum());
    t->ch = c;
    t->profile = profile;
    t->fft = fft;
    t->display_rate = display_rate;
    std::vector<int> lens;
    t->rate = plan_stages(b->in_rate, b->ch[c].dc_max_bw, lens);             // m_OutputRate of the channel's CDemodulator
    // CTestBench::Reset, gui/testbench.cpp:535-575 (m_DisplaySkipValue is a qint32: the quotient is truncated)
    CSDR_TRY(cutesdr_fft_set_params(fft, kTestFftSize, 0, 0.0, t->rate));
    t->skip_value = (int)(t->rate / (kTestFftSize * (double)display_rate));
    t->skip_counter = -2;
    CSDR_TRY(cutesdr_fft_reset(fft));
    CSDR_CK(cudaMalloc(&t->d_frame, kTestFftSize * sizeof(float2)));
    CSDR_CK(cudaMemsetAsync(t->d_frame, 0, kTestFftSize * sizeof(float2), b->st));
    CSDR_CK(cudaStreamSynchronize(b->st));
    b->tap_spectra.push_back(std::move(t));
    return CUTESDR_OK;
}

int cutesdr_bank_tap_spectrum_frames(cutesdr_bank* b, int c, int profile, long long* frames)
{
    if (!b || !frames) { set_error("tap_spectrum_frames: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    for (auto& t : b->tap_spectra)
        if (t->ch == c && t->profile == profile) { *frames = t->frames; return CUTESDR_OK; }
    set_error("tap_spectrum_frames: no spectrum attached to channel %d profile %d", c, profile);
    return CUTESDR_E_STATE;
}

int cutesdr_bank_tap_size(cutesdr_bank* b, int c, int profile, long* n_floats)
{
    if (!b || !n_floats || c < 0 || c >= b->nch || profile < 1 || profile > 4) { set_error("tap_size: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    *n_floats = (long)b->ch[c].tap[profile].size();
    return CUTESDR_OK;
}

int cutesdr_bank_tap_read(cutesdr_bank* b, int c, int profile, float* out, long cap_floats)
{
    if (!b || !out || c < 0 || c >= b->nch || profile < 1 || profile > 4) { set_error("tap_read: bad arguments"); return CUTESDR_E_ARG; }
    std::lock_guard<std::mutex> lk(b->mu);
    std::vector<float>& v = b->ch[c].tap[profile];
    long n = std::min<long>(cap_floats, (long)v.size());
    memcpy(out, v.data(), n * sizeof(float));
    v.clear();
    return (int)std::min<long>(n, 0x7fffffff);
}

}  // extern "C"
