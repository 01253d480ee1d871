/* cutesdr_cuda.h -- C ABI of libcutesdr_cuda: the B200 (sm_100a) implementation of
 * CuteSDR's receive DSP chain as a batched multi-channel receiver.
 *
 * The reference has no FFI layer: its dsp/ class headers ARE the API. Each entry point
 * below therefore names the reference method it replaces (file:line relative to the
 * reference tree). A maintainer binds them from C++ shims that keep the reference class
 * names -- see INTEGRATION.md and cutesdr_b200/compat/dsp/ (drop-in headers).
 *
 * Conventions
 *   - every function returns an int status: CUTESDR_OK (0) or a negative CUTESDR_E_*;
 *     the reference's own return value (sample counts, rates, flags) comes back through
 *     an out-parameter or, for `*_process`, as a non-negative return value.
 *   - sample scale is int16 full scale (+-32767), as every dB constant of the reference
 *     assumes (dsp/fft.cpp:19, dsp/agc.cpp:69, dsp/smeter.cpp:47).
 *   - complex buffers are interleaved re,im. `*_f32` entry points take complex64 (the fast
 *     path); the plain ones take the reference's TYPECPX = two doubles (dsp/datatypes.h:25-39).
 *   - handles are not thread-safe against concurrent `process` calls; setters may be called
 *     from another thread between blocks (they take the handle's mutex, like the reference's
 *     per-object QMutex) and take effect at the next DSP block.
 *   - there is no CPU fallback: without a CUDA device every create call fails with
 *     CUTESDR_E_CUDA.
 */
#ifndef CUTESDR_CUDA_H
#define CUTESDR_CUDA_H

#include <stdint.h>

#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

#define CUTESDR_OK 0
#define CUTESDR_E_ARG (-1)      /* bad argument / handle */
#define CUTESDR_E_CUDA (-2)     /* CUDA runtime error (see cutesdr_last_error) */
#define CUTESDR_E_STATE (-3)    /* call sequence error (e.g. process before set_demod) */
#define CUTESDR_E_NOMEM (-4)

/* demod modes, dsp/demodulator.h:20-28 */
#define CUTESDR_DEMOD_AM 0
#define CUTESDR_DEMOD_SAM 1
#define CUTESDR_DEMOD_FM 2
#define CUTESDR_DEMOD_USB 3
#define CUTESDR_DEMOD_LSB 4
#define CUTESDR_DEMOD_CWU 5
#define CUTESDR_DEMOD_CWL 6

/* tDemodInfo without its QString / click-resolution GUI fields, dsp/demodulator.h:35-54 */
typedef struct cutesdr_demod_info {
    int HiCut, HiCutmin, HiCutmax;
    int LowCut, LowCutmin, LowCutmax;
    int Offset;
    int SquelchValue;
    int AgcSlope, AgcThresh, AgcManualGain, AgcDecay;
    int AgcOn, AgcHangOn;
} cutesdr_demod_info;

const char* cutesdr_last_error(void);
const char* cutesdr_version(void);
int cutesdr_device_count(int* n);

/* Device microbenchmarks behind bench.py's roofline denominators (no reference counterpart; SURVEY.md 8d asks for
 * them): which = 0 FP32 FMA issue peak [TFLOP/s], 1 tcgen05 kind::tf32 dense peak [TFLOP/s], 2 tcgen05 kind::f16
 * dense peak [TFLOP/s], 3 HBM copy bandwidth [GB/s, read + write]. Each takes a few milliseconds. */
int cutesdr_microbench(int device, int which, double* value);
/* Page-locked host memory for the iq / audio buffers of the pipelined entry points (cutesdr_bank_process_async*), for
 * hosts that do not link the CUDA runtime themselves (no reference counterpart: the reference's buffers are plain
 * arrays, dsp/demodulator.cpp:77-80). */
int cutesdr_host_alloc(void** p, size_t bytes);
void cutesdr_host_free(void* p);
/* cudaMemGetInfo of `device` (tests use it to show that a long-running, constantly retuned bank does not grow) */
int cutesdr_device_memory(int device, long long* free_bytes, long long* total_bytes);

/* ======================================================================================
 * Receiver bank: N virtual receivers ( = N independent CDemodulator objects,
 * dsp/demodulator.h:56-100) fed from ONE wideband complex stream.
 * ====================================================================================== */
typedef struct cutesdr_bank cutesdr_bank;

/* N x { CDemodulator(); SetInputSampleRate(in_rate) }   dsp/demodulator.cpp:47-60,94-101 */
int cutesdr_bank_create(cutesdr_bank** out, int n_channels, double in_rate, int device);
void cutesdr_bank_destroy(cutesdr_bank* b);

/* CDemodulator::SetDemod(Mode, info) for channel ch          dsp/demodulator.cpp:107-157
 * On a running bank every change is live and touches this channel only: filter / AGC / squelch parameters are swapped
 * for the next burst; a mode change that alters the decimation chain (another m_DesiredMaxOutputBandwidth, :116-121)
 * restarts THIS channel's chain and demodulator from zero state -- it moves to a slot of the channel group with the new
 * chain -- while all other channels run on bit-identically. (Only a chain whose DSP block length differs from the
 * bank's forces a rebuild of the whole bank.) */
int cutesdr_bank_set_demod(cutesdr_bank* b, int ch, int mode, const cutesdr_demod_info* info);
/* CDemodulator::SetDemodFreq(Freq) for channel ch             dsp/demodulator.h:68-69 */
int cutesdr_bank_set_demod_freq(cutesdr_bank* b, int ch, double freq);
/* CDemodulator::GetOutputRate()                               dsp/demodulator.h:63 */
int cutesdr_bank_get_output_rate(cutesdr_bank* b, int ch, double* rate);
/* m_InBufLimit: input samples per DSP block                   dsp/demodulator.cpp:145-146 */
int cutesdr_bank_block_length(cutesdr_bank* b, int* n);
/* GetSMeterPeak()/GetSMeterAve(): either pointer may be NULL. Reading the peak resets the held
 * peak (CSMeter::GetPeak, dsp/smeter.cpp:99-104); a NULL `peak` leaves it alone (GetAve).   dsp/demodulator.h:64-65 */
int cutesdr_bank_get_smeter(cutesdr_bank* b, int ch, double* peak, double* ave);

/* CSdrInterface::SetupNoiseProc -> CNoiseProc::SetupBlanker on the shared wideband stream
 *                                                             dsp/noiseproc.cpp:77-119 */
int cutesdr_bank_set_noiseproc(cutesdr_bank* b, int on, double threshold, double width_us);

/* Optional per-channel CFractResampler to `audio_rate` after the demodulator, as
 * CSoundOut::PutOutQueue does (interface/soundout.cpp:204,262): Rate = OutputRate/audio_rate.
 * audio_rate <= 0 disables it (audio is delivered at the demodulator output rate). */
int cutesdr_bank_set_audio_rate(cutesdr_bank* b, double audio_rate);

/* Stereo output, CDemodulator::ProcessData(int, TYPECPX*, TYPECPX*) (dsp/demodulator.cpp:221-273): audio
 * rows then hold interleaved (left,right) float32 pairs, n_out counts frames and audio_stride (in floats)
 * must hold 2 floats per frame. AM and FM put the mono signal on both channels, SSB/CW pass the complex
 * filter output through, SAM splits lower/upper sideband with its Hilbert pair (dsp/samdemod.cpp:115-158).
 * Switching rebuilds the bank (all channel state restarts). With set_audio_rate the (left,right) frames go through the
 * stereo form of CFractResampler (dsp/fractresampler.cpp:194-249, call site interface/soundout.cpp:204). */
int cutesdr_bank_set_stereo(cutesdr_bank* b, int stereo);

/* N x CDemodulator::ProcessData(n_in, iq, audio) -- mono    dsp/demodulator.cpp:163-215
 * iq: n_in complex64 samples in HOST memory (any n_in; blocks are cut every block_length
 * samples exactly like m_pDemodInBuf). audio: HOST float32 [n_channels][audio_stride];
 * n_out[ch] receives the number of samples written for channel ch during this call
 * (0 or 1024 per completed DSP block without the resampler). Returns the maximum n_out. */
int cutesdr_bank_process(cutesdr_bank* b, int n_in, const float* iq, float* audio, int audio_stride, int* n_out);

/* Wire-format ingest (interface/netiobase.cpp:497-527): the same as cutesdr_bank_process / _async with
 * the samples still in the radio's integer format -- fmt 1 = interleaved little-endian int16 I,Q (4 bytes
 * per sample, value = the integer), fmt 2 = packed little-endian int24 I,Q (6 bytes per sample, value =
 * integer/256), fmt 0 = complex64. The unpack happens inside kernel 1's tile load, so H2D traffic is half
 * (int16) or three quarters (int24) of the float path and no conversion pass exists. These two take bare
 * samples; cutesdr_bank_process_packets below takes the radio's datagrams as they arrive. */
#define CUTESDR_FMT_CF32 0
#define CUTESDR_FMT_CS16 1
#define CUTESDR_FMT_CS24 2
int cutesdr_bank_process_raw(cutesdr_bank* b, int n_in, const void* data, int fmt, float* audio, int audio_stride, int* n_out);
int cutesdr_bank_process_async_raw(cutesdr_bank* b, int n_in, const void* data, int fmt, float* audio, int audio_stride, int* n_out);
/* The radio's UDP datagrams as received (CUdpThread::OnreadyRead, interface/netiobase.cpp:464-534): n_packets
 * datagrams of packet_bytes each, back to back in HOST memory; packet_bytes = 1028 (4-byte header + 256 int16 I/Q
 * pairs) or 1444 (header + 240 packed int24 pairs). Bytes 2-3 of the header are the little-endian sequence number;
 * gaps are counted exactly as the reference counts m_MissedPackets (:487-496) and read with
 * cutesdr_bank_missed_packets (reset != 0 clears the count and restarts the sequence, as a new CUdpThread does).
 * The payloads then take the cutesdr_bank_process_raw path (any packet count; DSP blocks are cut every block_length
 * samples like m_pDemodInBuf, CIQDataThread::run :571-600 -> ProcessIQData). */
#define CUTESDR_PKT_LENGTH_16 1028
#define CUTESDR_PKT_LENGTH_24 1444
int cutesdr_bank_process_packets(cutesdr_bank* b, int n_packets, const void* packets, int packet_bytes, float* audio,
                                 int audio_stride, int* n_out);
int cutesdr_bank_missed_packets(cutesdr_bank* b, long long* missed, int reset);
/* Pipelined form for a block that is already in DEVICE memory (multi-GPU: the NCCL broadcast of rank 0's block lands
 * there). d_iq (complex64[n_in]) is read in stream order with respect to src_stream (a cudaStream_t): the library
 * waits for the work queued on it so far and makes it wait until the block has been taken over, so the caller can
 * queue the next broadcast into the same buffer immediately. audio is a pinned HOST buffer as in process_async. */
int cutesdr_bank_process_async_device(cutesdr_bank* b, int n_in, const void* d_iq, void* src_stream, float* audio,
                                      int audio_stride, int* n_out);

/* ---- multi-GPU (one process per GPU). Channels shard across ranks -- every rank owns a bank with its
 * slice of the channel list -- and the only exchange is the wideband block: rank 0 receives it from the host and it is
 * broadcast with NCCL over NVLink / NVSwitch. No reference counterpart (the reference is single-threaded CPU code); the
 * per-rank calls are the same CDemodulator-shaped calls as above. NCCL is loaded at run time (libnccl.so.2).
 *   rank 0:     cutesdr_mgpu_unique_id(id)  -> hand the 128 bytes to the other processes (MPI, a socket, torch.distributed ...)
 *   every rank: cutesdr_mgpu_init(&m, id, rank, world, device); cutesdr_mgpu_channel_slice(n, rank, world, &first, &count);
 *               bank of `count` channels; then per DSP block cutesdr_bank_process_async_bcast(...). */
typedef struct cutesdr_mgpu cutesdr_mgpu;
int cutesdr_mgpu_unique_id(void* id128);
int cutesdr_mgpu_init(cutesdr_mgpu** out, const void* id128, int rank, int world, int device);
void cutesdr_mgpu_destroy(cutesdr_mgpu* m);
int cutesdr_mgpu_info(cutesdr_mgpu* m, int* rank, int* world, int* nccl_version, long long* blocks, long long* bytes_bcast);
int cutesdr_mgpu_channel_slice(int n_channels, int rank, int world, int* first, int* count);
/* cutesdr_bank_process_async_raw for one DSP block whose samples only rank 0 has: iq_rank0 (PINNED host memory, format
 * fmt, n_in == block_length; ignored on the other ranks) is copied to rank 0's GPU in 4 MiB chunks and every chunk is
 * broadcast as soon as it has landed (the copy of chunk k+1 overlaps the broadcast of chunk k), directly into the
 * bank's input slot on every rank; the block's kernels, the audio D2H and the host-buffer contract are those of
 * cutesdr_bank_process_async. Every rank must make the same sequence of calls. world == 1 degenerates to process_async. */
int cutesdr_bank_process_async_bcast(cutesdr_bank* b, cutesdr_mgpu* m, int n_in, const void* iq_rank0, int fmt, float* audio,
                                     int audio_stride, int* n_out);

/* Pipelined form of cutesdr_bank_process for exactly one DSP block per call (n_in == block_length,
 * iq and audio in PINNED host memory): the call only queues work -- the H2D copy of this block runs on
 * a copy stream under the previous block's kernels, the D2H of finished audio on another. n_out[] is
 * filled on return (burst timing is deterministic); the audio bytes are valid after
 * cutesdr_bank_synchronize. iq must stay unchanged until the second following call (or synchronize);
 * audio rows of a call that produced output must not be reused before they were synchronised. */
int cutesdr_bank_process_async(cutesdr_bank* b, int n_in, const float* iq, float* audio, int audio_stride, int* n_out);

/* Same, with the block already resident in device memory (d_iq: complex64[n_in], n_in must
 * equal block_length) and results left in device memory: d_audio float32
 * [n_channels][audio_stride] (may be NULL to keep them internal). Asynchronous on the bank's
 * stream; *n_out_max is known on return (burst timing is deterministic). d_iq must be 16-byte
 * aligned (kernel 1 reads it with vector loads); CUTESDR_E_ARG otherwise. */
int cutesdr_bank_process_device(cutesdr_bank* b, const void* d_iq, int n_in, void* d_audio, int audio_stride,
                                int* n_out_max);
/* same with the device block still in a wire format (fmt as in cutesdr_bank_process_raw) */
int cutesdr_bank_process_device_raw(cutesdr_bank* b, const void* d_data, int fmt, int n_in, void* d_audio, int audio_stride,
                                    int* n_out_max);
/* process_device queues work on the bank's main stream (decimation) and on internal side streams
 * (FIR / AGC / demodulator bursts, which overlap the next blocks' decimation). synchronize waits
 * for all of it; join only orders the main stream after all queued burst work, so that an event
 * recorded on the main stream afterwards covers everything (used for device timing). */
int cutesdr_bank_synchronize(cutesdr_bank* b);
int cutesdr_bank_join(cutesdr_bank* b);
/* Device pointer (complex64[n]) of the most recent DSP block as the channels saw it, i.e. after the
 * noise blanker -- the samples CSdrInterface::ProcessIQData hands to the display FFT
 * (interface/sdrinterface.cpp:884-909). Feed slices of it to cutesdr_fft_put_device. Valid until the
 * next process call. */
int cutesdr_bank_last_block(cutesdr_bank* b, const void** d_block, int* n);
/* cudaStream_t of the bank, as void* (for event timing on the launching stream) */
int cutesdr_bank_stream(cutesdr_bank* b, void** stream);
/* kernels launched by this bank so far */
int cutesdr_bank_launch_count(cutesdr_bank* b, long long* n);
/* CUDA-event timing of the dominant kernel on the bank's stream. kernel_timing(1) brackets every
 * launch of kernel `which` (0 = k_mix_cic, the fused NCO + CIC cascade) with two events;
 * kernel_time returns the accumulated milliseconds and launch count since the last read. */
int cutesdr_bank_kernel_timing(cutesdr_bank* b, int enable);
int cutesdr_bank_kernel_time(cutesdr_bank* b, int which, double* ms_total, long long* launches);
/* What bounds kernel 1 (which = 0; which = 1 asks instead whether every group runs the fp16-operand form of
 * kernel 1T, kind::f16): on_tensor_cores = 1 when every chain group runs kernel 1T (the NCO mix + first
 * four CIC3 stages of dsp/downconvert.cpp:186-263,425-460 as a tcgen05 GEMM); flops_per_block = the GEMM's
 * multiply-adds x 2 per DSP block, each real product counted once (the 3 tf32 partial products that emulate one
 * fp32 product are one). bench.py's roofline line is built from this and cutesdr_bank_kernel_time. */
int cutesdr_bank_kernel_model(cutesdr_bank* b, int which, int* on_tensor_cores, double* flops_per_block);

/* Test-bench taps (the reference's PROFILE_1..4 display taps, gui/testbench.cpp:71-81,
 * dsp/demodulator.cpp:175,180,187,208): 1 = after CDownConvert (complex), 2 = after CFastFIR
 * (complex), 3 = after CAgc (complex), 4 = demodulated audio (real). When enabled for a
 * channel the stream is accumulated on the host as float32 and read back with tap_read. */
int cutesdr_bank_tap_enable(cutesdr_bank* b, int ch, unsigned profile_mask);
int cutesdr_bank_tap_size(cutesdr_bank* b, int ch, int profile, long* n_floats);
int cutesdr_bank_tap_read(cutesdr_bank* b, int ch, int profile, float* out, long cap_floats);

/* The test bench's spectrum display of one tap, kept on the device (CTestBench::DisplayData frequency-domain branch,
 * gui/testbench.cpp:583-611; set-up CTestBench::Reset :535-575 and the constructor :128-132): the tap's samples are
 * appended to a TEST_FFTSIZE = 2048 sample frame in stream order and every m_DisplaySkipValue-th full frame
 * (m_DisplaySkipValue = (int)(rate / (2048 * display_rate)), counter starting at -2) goes through PutInDisplayFFT of
 * `fft`, which this call sets to SetFFTParams(2048, FALSE, 0.0, rate) + ResetFFT like the test bench does. Nothing
 * synchronises: read the result with cutesdr_fft_get_screen / get_plot whenever the GUI timer fires. Real taps
 * (PROFILE_4 mono) enter as (x, 0) like the TYPEREAL overload (:657-658); PROFILE_4 is the demodulator output at the
 * channel's output rate (in front of the bank resampler) and needs an audio destination in the process call.
 * fft == NULL detaches; detach before destroying the fft object. frames = PutInDisplayFFT calls made so far. */
struct cutesdr_fft;
int cutesdr_bank_tap_spectrum(cutesdr_bank* b, int ch, int profile, struct cutesdr_fft* fft, int display_rate);
int cutesdr_bank_tap_spectrum_frames(cutesdr_bank* b, int ch, int profile, long long* frames);

/* ======================================================================================
 * Single-object entry points (one reference object each; TYPECPX = double pairs)
 * ====================================================================================== */

/* ---- CDownConvert, dsp/downconvert.h:23-120 ---- */
typedef struct cutesdr_downconvert cutesdr_downconvert;
int cutesdr_downconvert_create(cutesdr_downconvert** out, int device);
void cutesdr_downconvert_destroy(cutesdr_downconvert* h);
/* SetFrequency                                                 dsp/downconvert.cpp:98-107 */
int cutesdr_downconvert_set_frequency(cutesdr_downconvert* h, double nco_freq);
/* SetCwOffset                                                  dsp/downconvert.h:29 */
int cutesdr_downconvert_set_cw_offset(cutesdr_downconvert* h, double offset);
/* SetDataRate -> *out_rate                                     dsp/downconvert.cpp:114-173 */
int cutesdr_downconvert_set_data_rate(cutesdr_downconvert* h, double in_rate, double max_bw, double* out_rate);
/* stage list as tap counts (3 = CIC3, 11, 15, ... 51) */
int cutesdr_downconvert_stages(cutesdr_downconvert* h, int* lens, int cap, int* n);
/* ProcessData(InLength, in, out): returns output count (>=0) or an error (<0). `in` is
 * NOT modified (the reference overwrites it with the mixer product; no caller reads that).
 *                                                              dsp/downconvert.cpp:186-263 */
int cutesdr_downconvert_process(cutesdr_downconvert* h, int n_in, const double* in, double* out);
int cutesdr_downconvert_process_f32(cutesdr_downconvert* h, int n_in, const float* in, float* out);

/* ---- CFastFIR, dsp/fastfir.h:19-44 ---- */
typedef struct cutesdr_fastfir cutesdr_fastfir;
int cutesdr_fastfir_create(cutesdr_fastfir** out, int device);
void cutesdr_fastfir_destroy(cutesdr_fastfir* h);
/* SetupParameters(FLoCut, FHiCut, Offset, SampleRate)         dsp/fastfir.cpp:178-259 */
int cutesdr_fastfir_setup(cutesdr_fastfir* h, double lo_cut, double hi_cut, double offset, double sample_rate);
/* ProcessData: returns 1024*k output samples (k>=0)            dsp/fastfir.cpp:268-306 */
int cutesdr_fastfir_process(cutesdr_fastfir* h, int n_in, const double* in, double* out);
int cutesdr_fastfir_process_f32(cutesdr_fastfir* h, int n_in, const float* in, float* out);

/* ---- CFft (display path), dsp/fft.h:24-85 ---- */
typedef struct cutesdr_fft cutesdr_fft;
int cutesdr_fft_create(cutesdr_fft** out, int device);
void cutesdr_fft_destroy(cutesdr_fft* h);
/* SetFFTParams(size, invert, dBCompensation, SampleFreq)      dsp/fft.cpp:118-243 */
int cutesdr_fft_set_params(cutesdr_fft* h, int size, int invert, double db_compensation, double sample_freq);
/* SetFFTAve / ResetFFT                                         dsp/fft.cpp:103-113,248-259 */
int cutesdr_fft_set_ave(cutesdr_fft* h, int ave);
int cutesdr_fft_reset(cutesdr_fft* h);
/* PutInDisplayFFT(n, InBuf) -> *total_count                   dsp/fft.cpp:267-288 */
int cutesdr_fft_put(cutesdr_fft* h, int n, const double* in, int* total_count);
int cutesdr_fft_put_f32(cutesdr_fft* h, int n, const float* in, int* total_count);
/* same, frame already in device memory (complex64[n]) -- e.g. a slice of the bank's wideband block */
int cutesdr_fft_put_device(cutesdr_fft* h, int n, const void* d_in, int* total_count);
/* same, stream-ordered and without host synchronisation (the concurrent spectrum of a running bank): the frame is
 * taken over in order with respect to src_stream (a cudaStream_t, e.g. cutesdr_bank_stream) -- it may be overwritten
 * as soon as the call returns --, the transform runs on the object's own stream, and the overload flag / averages are
 * picked up by the next GetScreenIntegerFFTData / get_plot, which waits for this object's work only */
int cutesdr_fft_put_device_async(cutesdr_fft* h, int n, const void* d_in, void* src_stream, int* total_count);
int cutesdr_fft_launch_count(cutesdr_fft* h, long long* n);
/* GetScreenIntegerFFTData(...) -> out[max_width], *overload   dsp/fft.cpp:308-410 */
int cutesdr_fft_get_screen(cutesdr_fft* h, int max_height, int max_width, double max_db, double min_db,
                           int start_freq, int stop_freq, int32_t* out, int* overload);
/* CPlotter::draw's two mappings of one spectrum (gui/plotter.cpp:429-456) from ONE pass: the 255-level waterfall row
 * (GetScreenIntegerFFTData(255, w, ...)) and the 2-D trace (GetScreenIntegerFFTData(h, w, ...)) */
int cutesdr_fft_get_plot(cutesdr_fft* h, int trace_height, int width, double max_db, double min_db, int start_freq,
                         int stop_freq, int32_t* waterfall_row, int32_t* trace, int* overload);
/* I/Q DC offset subtracted from the display copy of the samples before windowing and the overload test
 * (m_NCOSpurOffsetI/Q, interface/sdrinterface.cpp:889-894) */
int cutesdr_fft_set_dc_offset(cutesdr_fft* h, double off_i, double off_q);
/* FwdFFT / RevFFT(TYPECPX* pInOutBuf): in-place complex transform of the object's size, unnormalised.
 * As in the reference, "forward" is the e^{+j 2 pi nk/N} kernel and "reverse" its conjugate
 *                                                              dsp/fft.cpp:416-426 */
int cutesdr_fft_fwd(cutesdr_fft* h, double* io);
int cutesdr_fft_rev(cutesdr_fft* h, double* io);
/* averaged log-power buffer in bels, fft-shifted (index N/2 = DC) -- test tap */
int cutesdr_fft_get_ave(cutesdr_fft* h, float* out, int cap);

/* ---- CAgc (complex path), dsp/agc.h:18-63 ---- */
typedef struct cutesdr_agc cutesdr_agc;
int cutesdr_agc_create(cutesdr_agc** out, int device);
void cutesdr_agc_destroy(cutesdr_agc* h);
/* SetParameters                                                dsp/agc.cpp:104-167 */
int cutesdr_agc_set_parameters(cutesdr_agc* h, int agc_on, int use_hang, int threshold, int manual_gain,
                               int slope, int decay, double sample_rate);
/* ProcessData(Length, in, out) complex, in place allowed      dsp/agc.cpp:174-296 */
int cutesdr_agc_process(cutesdr_agc* h, int n, const double* in, double* out);

/* ---- CFractResampler, dsp/fractresampler.h:16-33 ---- */
typedef struct cutesdr_resampler cutesdr_resampler;
int cutesdr_resampler_create(cutesdr_resampler** out, int device);
void cutesdr_resampler_destroy(cutesdr_resampler* h);
/* Init(MaxInputSize)                                           dsp/fractresampler.cpp:85-135 */
int cutesdr_resampler_init(cutesdr_resampler* h, int max_input_size);
/* Resample overloads: real->real, cpx->cpx, real->int16 (gain, clip), cpx->stereo int16
 * return output count                                          dsp/fractresampler.cpp:144-352 */
int cutesdr_resampler_real(cutesdr_resampler* h, int n, double rate, const double* in, double* out);
int cutesdr_resampler_cpx(cutesdr_resampler* h, int n, double rate, const double* in, double* out);
int cutesdr_resampler_mono16(cutesdr_resampler* h, int n, double rate, const double* in, int16_t* out, double gain);
int cutesdr_resampler_stereo16(cutesdr_resampler* h, int n, double rate, const double* in, int16_t* out, double gain);

/* ---- CIir, dsp/iir.h:16-40 (one direct-form-2 biquad; embedded in CSdrInterface, interface/sdrinterface.h:178) ---- */
#define CUTESDR_IIR_LP 0
#define CUTESDR_IIR_HP 1
#define CUTESDR_IIR_BP 2
#define CUTESDR_IIR_BR 3
typedef struct cutesdr_iir cutesdr_iir;
/* CIir(): InitBR(25000, 1000, 100000)                          dsp/iir.cpp:77-80 */
int cutesdr_iir_create(cutesdr_iir** out, int device);
void cutesdr_iir_destroy(cutesdr_iir* h);
/* InitLP / InitHP / InitBP / InitBR(F0Freq, FilterQ, SampleRate); clears the delay storage   dsp/iir.cpp:86-163 */
int cutesdr_iir_init(cutesdr_iir* h, int kind, double f0, double q, double sample_rate);
/* ProcessFilter(InLength, in, out) real / complex (interleaved doubles); returns n           dsp/iir.cpp:169-201 */
int cutesdr_iir_process_real(cutesdr_iir* h, int n, const double* in, double* out);
int cutesdr_iir_process_cpx(cutesdr_iir* h, int n, const double* in, double* out);

/* ---- CNoiseProc, dsp/noiseproc.h:23-53 ---- */
typedef struct cutesdr_noiseproc cutesdr_noiseproc;
int cutesdr_noiseproc_create(cutesdr_noiseproc** out, int device);
void cutesdr_noiseproc_destroy(cutesdr_noiseproc* h);
/* SetupBlanker(On, Threshold, Width, SampleRate)              dsp/noiseproc.cpp:77-119 */
int cutesdr_noiseproc_setup(cutesdr_noiseproc* h, int on, double threshold, double width_us, double sample_rate);
/* ProcessBlanker(n, in, out) in place allowed                  dsp/noiseproc.cpp:121-176 */
int cutesdr_noiseproc_process(cutesdr_noiseproc* h, int n, const double* in, double* out);
int cutesdr_noiseproc_process_f32(cutesdr_noiseproc* h, int n, const float* in, float* out);

/* ---- CDemodulator (one receiver; a 1-channel bank), dsp/demodulator.h:56-100 ---- */
typedef struct cutesdr_demodulator cutesdr_demodulator;
int cutesdr_demodulator_create(cutesdr_demodulator** out, int device);
void cutesdr_demodulator_destroy(cutesdr_demodulator* h);
int cutesdr_demodulator_set_input_sample_rate(cutesdr_demodulator* h, double rate);   /* dsp/demodulator.cpp:94-101 */
int cutesdr_demodulator_set_demod(cutesdr_demodulator* h, int mode, const cutesdr_demod_info* info); /* :107-157 */
int cutesdr_demodulator_set_demod_freq(cutesdr_demodulator* h, double freq);          /* dsp/demodulator.h:68-69 */
int cutesdr_demodulator_get_output_rate(cutesdr_demodulator* h, double* rate);
int cutesdr_demodulator_get_smeter(cutesdr_demodulator* h, double* peak, double* ave);
/* ProcessData(InLength, TYPECPX* in, TYPEREAL* out) mono; returns samples written to out
 *                                                              dsp/demodulator.cpp:163-215 */
int cutesdr_demodulator_process(cutesdr_demodulator* h, int n_in, const double* in, double* out);
/* ProcessData(InLength, TYPECPX* in, TYPECPX* out) stereo; returns frames written (2 doubles each). Mixing
 * mono and stereo calls on one object restarts its state (the reference keeps one set of demod objects).
 *                                                              dsp/demodulator.cpp:221-273 */
int cutesdr_demodulator_process_stereo(cutesdr_demodulator* h, int n_in, const double* in, double* out);

#ifdef __cplusplus
}
#endif
#endif /* CUTESDR_CUDA_H */
