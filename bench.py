#!/usr/bin/env python3
"""bench.py -- headline benchmark of the receive-chain bank: input Msps x channels per B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg4|cfg5|cfg5q|cfg3|cfg1|cfg2] [--ingest cf32|cs16] [--scaling weak|strong]

One STEP = 0.2 s of the wideband stream = 20 DSP blocks (every config is defined on 0.2 s of signal; one
DSP block is m_InBufLimit samples, ~10 ms) pushed through the whole chain -- NCO mix + CIC/half-band cascade,
overlap-save FIR, S-meter, AGC, demodulator (+ resampler; + noise blanker and the concurrent 65536-point spectrum at
10 frames/s for config 5) -- for every channel of the rank's bank. Default workload: BASELINE config 4, 1024-channel
NBFM + CFractResampler to 48 kHz on the "100 Msps" stream (100 147 200 sps), 1024 channels PER GPU (weak
scaling: every GPU sees the same wideband stream and owns its own channel slice; --scaling strong splits the
config's channel count over the GPUs instead).

  value   : device-timed (CUDA events on the bank's stream), the stream's 20 distinct blocks resident in HBM (160 MB
            of complex64, more than the 126 MB L2) and cycled.
  e2e     : the same through the C ABI with HOST buffers: cutesdr_bank_process_async[_raw] (N = 1) or
            cutesdr_bank_process_async_bcast (N > 1: rank 0's H2D in chunks + NCCL broadcast inside the library); the
            H2D of every block and the D2H of every channel's audio are inside the timed region.
  roofline: kernel 1, timed with CUDA events around every launch on the bank's stream. Tensor-core path (k_mix_tc):
            bound "tensor", algorithmic GEMM flops / launch time vs MEASURED_PEAKS.json's dense bf16 figure (burst when the
            SM clock stayed at its maximum through the timed region, else sustained), halved for kind::tf32; the
            tcgen05 / FP32-FMA peaks measured live by cutesdr_microbench ride along. CUDA-core path (k_mix_cic): bound
            "fp32": algorithmic FP32 ops / launch time vs the measured FMA issue peak. The SURVEY 8(d)
            per-channel streaming model (8 B per sample*channel vs HBM) is reported as hbm_model, never as the bound.
  cpu_baseline / --impl reference : the UNMODIFIED reference dsp/*.cpp (oracle/_ref) on all host cores, persistent
            CDemodulator (+CFractResampler) objects built OUTSIDE the timed region, one step = the same 0.2 s of the same
            stream over a bounded channel sample.
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cutesdr_b200 import modes as M  # noqa: E402

BLOCKS_PER_STEP = 20


def _mixed(c):
    return (M.DEMOD_AM, M.DEMOD_SAM, M.DEMOD_FM, M.DEMOD_USB)[c % 4]


WORKLOADS = {
    "cfg4": dict(in_rate=100147200.0, nch=1024, pick=lambda c: M.DEMOD_FM, spacing=78125.0, audio_rate=48000.0, decim=2048,
                 desc="cfg4: 1024-ch NBFM (+LP biquad, CFractResampler->48 kHz) on 100.1472 Msps"),
    "cfg5": dict(in_rate=200294400.0, nch=4096, pick=_mixed, spacing=39000.0, audio_rate=0.0, decim=8192, blanker=True,
                 spectrum=True, impulses=20,
                 desc="cfg5: 4096-ch mixed AM/SAM/FM/USB with AGC (hang on every 8th), noise blanker (50, 50 us) and a concurrent "
                      "65536-pt spectrum (ave 4, 10 frames/s) on 200.2944 Msps"),
    "cfg5q": dict(in_rate=200294400.0, nch=1024, pick=_mixed, spacing=39000.0, audio_rate=0.0, decim=8192, blanker=True,
                  spectrum=True, impulses=20,
                  desc="cfg5 quarter: 1024-ch mixed AM/SAM/FM/USB with AGC, noise blanker and concurrent 65536-pt spectrum on 200.2944 Msps"),
    "cfg3": dict(in_rate=20000000.0, nch=256, pick=lambda c: M.DEMOD_USB if c % 2 == 0 else M.DEMOD_LSB, spacing=62500.0,
                 audio_rate=0.0, decim=512, desc="cfg3: 256-ch USB/LSB SSB bank (CDownConvert + CFastFIR + CSsbDemod + CAgc) on 20 Msps"),
    "cfg1": dict(in_rate=2000000.0, nch=1, pick=lambda c: M.DEMOD_AM, spacing=0.0, audio_rate=48000.0, decim=64, carrier0=250000.0,
                 desc="cfg1: single-channel AM (10 kHz BW, 48 kHz audio) on 2.0 Msps"),
}


def demod_info_for(m, c):
    if m == M.DEMOD_USB:
        return M.demod_info(m, HiCut=2800, LowCut=100, AgcHangOn=(c % 8 == 3))
    if m == M.DEMOD_LSB:
        return M.demod_info(m, HiCut=-100, LowCut=-2800)
    return M.demod_info(m, AgcHangOn=(c % 8 == 0))


def channel_plan(name, rank, world, scaling):
    """Global channel list of the job and this rank's slice of it."""
    from cutesdr_b200.dsp import channel_slice
    w = WORKLOADS[name]
    total = w["nch"] * (world if scaling == "weak" else 1)
    first, count = channel_slice(total, rank, world)
    modes = [w["pick"](c) for c in range(total)]
    spacing = w["spacing"]
    if total > 1 and spacing * total > 0.95 * w["in_rate"]:
        spacing = 0.95 * w["in_rate"] / total          # weak scaling at N > 1: keep every carrier inside the stream's band
    carriers = (np.arange(total) - total / 2 + 0.5) * spacing if total > 1 else np.array([w.get("carrier0", 0.0)])
    infos = [demod_info_for(m, c) for c, m in enumerate(modes)]
    return w, total, first, count, modes, carriers, infos


def block_length(in_rate):
    return int(in_rate / 100) & ~0xFF          # m_InBufLimit, dsp/demodulator.cpp:145-146


def make_stream(w, modes, carriers, ingest):
    """The job's 0.2 s of SYN-IQ: every channel's carrier is present. cf32: A = 16000/sqrt(Nch); cs16: a
    quarter of that, so the sum of a thousand carriers stays inside int16."""
    from cutesdr_b200.synth import syn_iq_fft
    L = block_length(w["in_rate"])
    n = BLOCKS_PER_STEP * L
    amp = 16000.0 if ingest == "cf32" else 4000.0
    iq, snapped = syn_iq_fft(w["in_rate"], n, modes, carriers, seed=20260 + len(modes), decim=w["decim"], total_amp=amp,
                             impulses=w.get("impulses", 0))
    return iq, snapped, L


def common_config(name, w, world, scaling, ingest, L, total):
    """`config` of the JSON line: identical keys and values in both arms."""
    return {"workload": w["desc"], "in_rate_sps": w["in_rate"], "channels_total": total,
            "channels_per_gpu": total // world, "block_length": L, "blocks_per_step": BLOCKS_PER_STEP,
            "samples_per_step": BLOCKS_PER_STEP * L, "ingest": ingest,
            "l2": "cycles %d distinct %d-sample blocks (%.0f MB > 126 MB L2)" % (BLOCKS_PER_STEP, L, BLOCKS_PER_STEP * L * (8 if ingest == "cf32" else 4) / 1e6),
            "parallelism": "channels sharded over %d GPU(s), wideband block broadcast from rank 0" % world if world > 1 else "1 GPU"}


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock / power / throttle reasons sampled through NVML every ~2 ms DURING the timed region (nvidia-smi's
    100 ms loop cannot see a sub-second region); one more sample is taken right at stop so the record is never empty."""

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.stop_flag = False
        self.thr = None
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = None
            try:        # CUDA_VISIBLE_DEVICES may renumber the devices: find the NVML handle by UUID
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = None
            if self.h is None:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            self.nv = None

    def _sample(self):
        nv = self.nv
        try:
            sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
            self.rows.append((sm, reasons, pw))
        except Exception:
            pass

    def _loop(self):
        while not self.stop_flag:
            self._sample()
            time.sleep(0.002)

    def start(self):
        if self.nv is None:
            return
        self.thr = threading.Thread(target=self._loop, daemon=True)
        self.thr.start()

    def stop(self):
        if self.nv is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"], "samples": 0}
        self.stop_flag = True
        if self.thr:
            self.thr.join(timeout=1.0)
        self._sample()
        nv = self.nv
        try:
            mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception:
            mx = None
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                "hw_power_brake_slowdown": 0x80}
        seen = 0
        for r in self.rows:
            seen |= int(r[1])
        reasons = [k for k, b in bits.items() if seen & b]
        sm = [r[0] for r in self.rows]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": float(min(sm)) if sm else None,
                "sm_max_mhz": float(mx) if mx else None, "power_w_max": max((r[2] for r in self.rows), default=None),
                "reasons": reasons, "samples": len(sm), "how": "NVML, 2 ms period, inside the timed region"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))
        except Exception:
            pass
    return None


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------
class CpuReference:
    """The reference's own CPU implementation (oracle/_ref = the unmodified dsp/*.cpp) on all host cores: persistent
    CDemodulator (+CFractResampler) objects, channels spread over std::threads, every step = the same 0.2 s stream."""

    def __init__(self, name, ch_per_core=1):
        from oracle import ref_binding as rb
        self.rb = rb
        w, total, first, count, modes, carriers, infos = channel_plan(name, 0, 1, "weak")
        self.w, self.total = w, total
        self.big = w["in_rate"] > 30e6
        if not rb.ref_available(big=self.big):
            raise RuntimeError("oracle/_ref not built")
        self.cores = os.cpu_count() or 1
        self.iq, snapped, self.L = make_stream(w, modes, carriers, "cf32")
        n_cpu = min(total, self.cores * int(os.environ.get("CUTESDR_BENCH_CPU_CH_PER_CORE", ch_per_core)))
        self.sel = [int(v) for v in np.linspace(0, total - 1, n_cpu).astype(int)]
        t0 = time.perf_counter()
        self.chains = rb.RefChainSet([modes[c] for c in self.sel], [-snapped[c] for c in self.sel], [infos[c] for c in self.sel],
                                     w["in_rate"], audio_rate=w["audio_rate"], keep_output=False, big=self.big)
        self.nb = self.fft = None
        if w.get("blanker"):
            self.nb = rb.RefNoiseProc(big=self.big)
            self.nb.SetupBlanker(True, 50.0, 50.0, w["in_rate"])
        if w.get("spectrum"):
            self.fft = rb.RefFft(big=self.big)
            self.fft.SetFFTParams(65536, False, 0.0, w["in_rate"])
            self.fft.SetFFTAve(4)
        self.setup_s = time.perf_counter() - t0

    def step(self):
        """one step; returns (seconds the whole job would take on these cores, raw seconds measured)"""
        iq = self.iq
        shared = 0.0
        if self.nb is not None:
            iq = self.iq.copy()
            shared += self.rb.blank_stream_f32(self.nb, iq)
        if self.fft is not None:
            t0 = time.perf_counter()
            for k in (0, BLOCKS_PER_STEP // 2):            # 10 frames/s
                self.fft.PutInDisplayFFT(iq[k * self.L + 100000:k * self.L + 100000 + 65536].astype(np.complex128))
                self.fft.GetScreenIntegerFFTData(255, 1024, 0.0, -140.0, int(-self.w["in_rate"] / 2), int(self.w["in_rate"] / 2))
            shared += time.perf_counter() - t0
        t = self.chains.run(iq, self.cores)
        # the per-stream work (blanker, spectrum) is paid once per stream, the chains scale with the channel count
        return shared + t * (self.total / len(self.sel)), shared + t

    def describe(self, steps):
        return ("%d of %d channels x %d samples (0.2 s of stream, %d blocks) per step, %d steps, %d threads; objects and "
                "CFractResampler::Init built outside the timed region (%.2f s); channel time scaled to the full bank%s"
                % (len(self.sel), self.total, len(self.iq), BLOCKS_PER_STEP, steps, self.cores, self.setup_s,
                   ", blanker + spectrum counted once per stream" if self.nb is not None else ""))


def cpu_reference_run(name, steps, warmup):
    ref = CpuReference(name)
    full, raw = [], []
    for it in range(warmup + steps):
        f, r = ref.step()
        if it >= warmup:
            full.append(f)
            raw.append(r)
    sec = float(np.sum(full))
    value = (len(ref.iq) * ref.total * steps) / sec / 1e6
    return ref, {"value": value, "cores": ref.cores, "kind": "reference", "ms_per_step": 1e3 * sec / steps,
                 "measured_s_per_step": float(np.mean(raw)), "sample": ref.describe(steps)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, args.steps)
    warm = max(1, min(args.warmup, 3))
    # bounded: the whole run stays within a few minutes whatever K the driver asks for
    budget_steps = int(os.environ.get("CUTESDR_BENCH_REF_MAX_STEPS", "24"))
    steps_run = min(steps, budget_steps)
    ref, r = cpu_reference_run(args.workload, steps_run, warm)
    w = ref.w
    cfg = common_config(args.workload, w, args.gpus, args.scaling, args.ingest, ref.L, ref.total * (args.gpus if args.scaling == "weak" else 1))
    line = {"impl": "reference", "metric": "input_msps_x_channels", "value": r["value"], "unit": "Msps*ch", "n_gpus": args.gpus,
            "steps": steps_run, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": r["value"], "unit": "Msps*ch", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "Msps*ch", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference CPU chain (unmodified dsp/*.cpp compiled headless, oracle/_ref), all host cores; value = samples x "
                    "channels of the full bank / time the full bank would take at the measured per-channel rate"}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# display FFT (config 2)
# ------------------------------------------------------------------------------------------------
def run_cfg2(args):
    """4096-point display FFT with window, averaging and integer screen mapping on 2.0 Msps IQ: one step = 256 frames
    through CFft::PutInDisplayFFT + GetScreenIntegerFFTData(255, 800, ...), each frame from a host buffer."""
    import cutesdr_b200 as cs
    from cutesdr_b200.synth import carrier_grid, syn_iq
    fs, N, frames = 2e6, 4096, 256
    modes = [M.DEMOD_AM] * 16
    iq = syn_iq(fs, N * frames, modes, carrier_grid(16, 100e3), seed=20262)
    cfg = {"workload": "cfg2: 4096-point display FFT with window, averaging (4) and integer screen mapping on 2.0 Msps IQ",
           "in_rate_sps": fs, "fft_size": N, "frames_per_step": frames, "samples_per_step": N * frames, "ingest": "cf32",
           "l2": "n/a (one 32 KB frame per call, host buffers)", "parallelism": "1 GPU"}
    if args.impl == "reference":
        from oracle import ref_binding as rb
        f = rb.RefFft()
        f.SetFFTParams(N, False, 0.0, fs)
        f.SetFFTAve(4)
        x = iq.astype(np.complex128)

        def step():
            for k in range(frames):
                f.PutInDisplayFFT(x[k * N:(k + 1) * N])
                f.GetScreenIntegerFFTData(255, 800, 0.0, -140.0, -1000000, 1000000)
        launches = 0
    else:
        f = cs.CFft()
        f.SetFFTParams(N, False, 0.0, fs)
        f.SetFFTAve(4)
        x = iq.astype(np.complex128)

        def step():
            for k in range(frames):
                f.PutInDisplayFFT(x[k * N:(k + 1) * N])
                f.GetScreenIntegerFFTData(255, 800, 0.0, -140.0, -1000000, 1000000)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    sec = time.perf_counter() - t0
    value = N * frames * args.steps / sec / 1e6
    line = {"metric": "input_msps_x_channels", "value": value, "unit": "Msps*ch", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.impl == "ours" else "f64", "data": "synthetic", "config": cfg,
            "e2e": {"value": value, "unit": "Msps*ch", "h2d_bytes_per_step": 8 * N * frames if args.impl == "ours" else 0,
                    "d2h_bytes_per_step": 4 * 800 * frames if args.impl == "ours" else 0},
            "frames_per_s": frames * args.steps / sec, "realtime_factor": value / (fs / 1e6),
            "note": "host-timed through the C ABI (every frame: H2D 32 KB, one fused kernel, screen kernel, D2H 3.2 KB, two syncs): "
                    "latency-bound by design, a display refreshes at 10-50 frames/s"}
    if args.impl == "reference":
        line["impl"] = "reference"
        line["cpu_baseline"] = {"value": value, "unit": "Msps*ch", "cores": 1, "kind": "reference", "sample": "%d frames per step, 1 thread" % frames}
        line["gpu_launches"] = 0
    else:
        n = C.c_longlong()
        f.L.cutesdr_fft_launch_count(f.h, C.byref(n))
        line["gpu_launches"] = int(n.value)
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def fp32_ops_per_sample(lens):
    """NCO rotate 4 + complex mix 4 + sum over stages of cost / 2^s (CIC3 3, N-tap half-band (N+3)/2)."""
    ops, div = 8.0, 1.0
    for n in lens:
        ops += (3.0 if n == 3 else (n + 3) / 2.0) / div
        div *= 2.0
    return ops


def run_ours(args):
    import torch
    import torch.distributed as dist
    import cutesdr_b200 as cs
    from cutesdr_b200.sharding import open_multi_gpu

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    name = args.workload
    w, total, first, nch, modes, carriers, infos = channel_plan(name, rank, world, args.scaling)
    mg = open_multi_gpu(rank, world, local)          # cutesdr_mgpu_*: the NCCL communicator lives inside the library

    def measure(ingest):
        """value (device-resident blocks), e2e (host blocks through the C ABI) and kernel 1's roofline for one wire format"""
        iq, snapped, L = make_stream(w, modes, carriers, ingest)
        fmt = 0 if ingest == "cf32" else 1
        bank = cs.ReceiverBank(nch, w["in_rate"], device=local)
        if w["audio_rate"] > 0:
            bank.SetAudioRate(w["audio_rate"])
        if w.get("blanker"):
            bank.SetupNoiseProc(True, 50.0, 50.0)
        for i in range(nch):
            c = first + i
            bank.SetDemod(i, modes[c], infos[c])
            bank.SetDemodFreq(i, -snapped[c])
        assert bank.block_length() == L
        fft = None
        if w.get("spectrum") and rank == 0:
            fft = cs.CFft(device=local)
            fft.SetFFTParams(65536, False, 0.0, w["in_rate"])
            fft.SetFFTAve(4)
        stream = torch.cuda.ExternalStream(bank.stream(), device=dev)
        nblk = BLOCKS_PER_STEP

        # the 0.2 s stream: pinned on the host, resident on the device
        if fmt == 0:
            host = torch.from_numpy(iq.view(np.float32).reshape(nblk, 2 * L)).pin_memory()
        else:
            i16 = np.empty((nblk, 2 * L), dtype=np.int16)
            v = iq.view(np.float32).reshape(nblk, 2 * L)
            np.clip(np.rint(v), -32767, 32767, out=v)
            i16[:] = v
            host = torch.from_numpy(i16).pin_memory()
        del iq
        d_blocks = host.to(dev)
        sample_bytes = 8 if fmt == 0 else 4
        audio_stride = 2304
        d_audio = torch.zeros((nch, audio_stride), dtype=torch.float32, device=dev)
        h_audio2 = [torch.zeros((nch, audio_stride), dtype=torch.float32).pin_memory() for _ in range(2)]
        n_out = np.zeros(nch, dtype=np.int32)
        spec_args = (600, 1024, 0.0, -140.0, int(-w["in_rate"] / 2), int(w["in_rate"] / 2))
        bank_stream = bank.stream()

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        def spectrum(k):
            # the concurrent 65536-point spectrum at 10 frames/s: one frame every 10 blocks, taken from the block the
            # channels saw (after the blanker), plus both plot mappings of CPlotter::draw
            if fft is not None and k % (nblk // 2) == 0:
                ptr, n = bank.last_block()
                fft.put_device_async(ptr + 8 * 100000, 65536, bank_stream)
                fft.GetPlot(*spec_args)

        def step_device():
            got = 0
            for k in range(nblk):
                got += bank.process_device(d_blocks[k].data_ptr(), L, d_audio.data_ptr(), audio_stride, fmt=fmt)
                spectrum(k)
            return got

        # ---- device-timed value
        for _ in range(args.warmup):
            step_device()
        barrier()
        bank.kernel_timing(True)
        bank.kernel_time(0)
        launches0 = bank.launch_count()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            for _ in range(args.steps):
                step_device()
            bank.join()             # main stream waits for the burst side streams: ev1 covers all work
            ev1.record(stream)
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        ms = ev0.elapsed_time(ev1)
        launches = bank.launch_count() - launches0
        k1_ms, k1_n = bank.kernel_time(0)
        bank.kernel_timing(False)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_max = float(t.item())
        samples_per_step = nblk * L
        value = (samples_per_step * total * args.steps) / (ms_max * 1e-3) / 1e6

        # ---- end to end through the C ABI: host block -> (rank 0 H2D -> NCCL broadcast inside the library) -> bank -> D2H audio
        e2e_steps = max(1, min(args.steps, 10))
        d2h = [0]

        def step_e2e():
            for k in range(nblk):
                hp = host[k].data_ptr()
                if world == 1:
                    m = bank.process_async_raw_ptr(L, hp, fmt, h_audio2[k & 1].data_ptr(), audio_stride, n_out)
                else:
                    m = bank.process_async_bcast_ptr(mg, L, hp if rank == 0 else None, fmt, h_audio2[k & 1].data_ptr(), audio_stride, n_out)
                d2h[0] += int(m) * nch * 4
                spectrum(k)

        step_e2e()
        bank.synchronize()
        barrier()
        d2h[0] = 0
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        bank.synchronize()
        torch.cuda.synchronize()
        t_e2e = time.perf_counter() - t0
        te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_value = (samples_per_step * total * e2e_steps) / float(te.item()) / 1e6
        mg_info = mg.info()

        if rank == 0:
            peaks = measured_peaks()
            hbm_peak = float(peaks["hbm_gbs"]) if peaks else 6650.0
            hbm_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
            live = {}
            for key, which in (("fp32_fma_tflops", 0), ("tcgen05_tf32_tflops", 1), ("tcgen05_f16_tflops", 2), ("hbm_copy_gbs", 3)):
                try:
                    live[key] = cs.microbench(which, local)
                except Exception as e:  # noqa: BLE001
                    live[key] = None
                    live[key + "_error"] = repr(e)
            roof = None
            if k1_n > 0 and k1_ms > 0:
                per_launch_s = (k1_ms / k1_n) * 1e-3
                groups = max(1, round(k1_n / max(1, args.steps * nblk)))
                units = float(L) * nch / groups                     # sample*channels one launch processes
                ach = 8.0 * units / per_launch_s / 1e9
                traffic = None
                try:
                    traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))[name]["dram_bytes_per_launch"]
                except Exception:
                    pass
                hbm_model = {"achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "peak_source": hbm_src,
                             "note": "SURVEY 8(d) per-channel streaming MODEL (8 B per sample*channel); not a DRAM figure and not the "
                                     "bound -- all channels share the staged samples; see traffic for the measured DRAM bytes"}
                on_tc, flops_block = bank.kernel_model(0)
                at_max = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
                if on_tc and flops_block > 0:
                    f16 = bank.kernel_model(1)[0]          # which = 1: fp16 hi/lo operand form in use?
                    if peaks:
                        dense = float(peaks["bf16_tflops"] if at_max else peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"]))
                        psrc = "MEASURED_PEAKS.json dense bf16 %s figure%s" % ("burst" if at_max else "sustained", "" if f16 else " / 2 (kind::tf32)")
                    else:
                        dense, psrc = 2250.0, "fallback: nominal 2.25 PFLOP/s bf16%s (B200_PROFILING.md)" % ("" if f16 else " / 2")
                    tpeak = dense if f16 else dense / 2.0
                    flops_launch = flops_block / groups
                    ach_t = flops_launch / per_launch_s / 1e12
                    roof = {"bound": "tensor", "achieved": ach_t, "peak": tpeak, "unit": "TFLOP/s", "frac": ach_t / tpeak, "traffic": traffic,
                            "kernel": "k_mix_tc (%s operands)" % ("fp16 hi/lo" if f16 else "tf32 hi/lo"), "launch_ms": k1_ms / k1_n,
                            "launches": k1_n, "peak_source": psrc + ("; SM clock at max through the timed region" if at_max else ""),
                            "flops_per_launch": flops_launch, "units_per_launch": units, "flops_per_unit": flops_launch / units,
                            "executed": {"tflops": (4.0 if f16 else 3.0) * ach_t, "frac": (4.0 if f16 else 3.0) * ach_t / tpeak,
                                         "note": ("int16 samples split exactly into fp16 hi + lo, coefficients into fp16 hi + lo: all four "
                                                  "partial-product MMAs per product, the tensor pipe executes 4x the algorithmic flops" if f16 else
                                                  "every fp32 product is three hi/lo partial-product MMAs: the tensor pipe executes 3x the algorithmic flops")},
                            "peaks_measured_live": live, "hbm_model": hbm_model}
                else:
                    ops = fp32_ops_per_sample(bank_stage_list(cs, w, modes[first], infos[first]))
                    fpeak = live.get("fp32_fma_tflops") or 2.0 * 148 * 128 * 1.965e9 / 1e12
                    ach_f = 2.0 * ops * units / per_launch_s / 1e12       # FMA = 2 flop
                    roof = {"bound": "fp32", "achieved": ach_f, "peak": fpeak, "unit": "TFLOP/s", "frac": ach_f / fpeak, "traffic": traffic,
                            "kernel": "k_mix_cic", "launch_ms": k1_ms / k1_n, "launches": k1_n,
                            "peak_source": "FP32 FMA issue peak measured live (cutesdr_microbench 0), FMA = 2 flop",
                            "ops_per_unit": ops, "units_per_launch": units,
                            "note": "algorithmic FP32 ops per sample*channel of the whole decimation ladder (SURVEY 8d) attributed to kernel 1 "
                                    "(it executes the NCO, the mixer and every CIC3 + the first half-band: > 90 % of them)",
                            "peaks_measured_live": live, "hbm_model": hbm_model}
        res = {"ingest": ingest, "L": L, "value": value, "ms_max": ms_max, "e2e_value": e2e_value, "e2e_steps": e2e_steps,
               "h2d": sample_bytes * L * nblk, "d2h": d2h[0] // e2e_steps, "launches": int(launches), "clocks": clocks,
               "roof": roof if rank == 0 else None, "mg_info": mg_info, "nblk": nblk}
        del bank
        torch.cuda.synchronize()
        return res

    nblk = BLOCKS_PER_STEP
    main_res = measure(args.ingest)
    alt_res = None
    if name == "cfg4" and not args.no_alt_ingest:
        alt_res = measure("cs16" if args.ingest == "cf32" else "cf32")
    L, value, ms_max, e2e_value, e2e_steps = main_res["L"], main_res["value"], main_res["ms_max"], main_res["e2e_value"], main_res["e2e_steps"]
    launches, clocks, roof, mg_info = main_res["launches"], main_res["clocks"], main_res["roof"], main_res["mg_info"]
    sample_bytes_step, d2h_step = main_res["h2d"], main_res["d2h"]
    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline:
            try:
                _, r = cpu_reference_run(name, steps=1, warmup=1)
                cpu = {"value": r["value"], "unit": "Msps*ch", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": "Msps*ch", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}
        cfg = common_config(name, w, world, args.scaling, args.ingest, L, total)
        line = {"metric": "input_msps_x_channels", "value": value, "unit": "Msps*ch", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": args.scaling,
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "realtime_factor": value / (w["in_rate"] * total / 1e6), "ms_per_block": ms_max / args.steps / nblk,
                "arithmetic": (("float32 chain; kernel 1 on tensor cores = int16 samples as exact fp16 hi + lo, every product as 4 hi/lo MMAs, fp32 accumulation"
                                if "fp16" in roof.get("kernel", "") else
                                "float32 chain; kernel 1 on tensor cores = every fp32 product as 3 hi/lo MMAs, fp32 accumulation")
                               if (roof and roof.get("bound") == "tensor") else "float32 chain (CUDA cores)"),
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Msps*ch", "h2d_bytes_per_step": sample_bytes_step,
                        "d2h_bytes_per_step": d2h_step, "steps": e2e_steps,
                        "api": "cutesdr_bank_process_async_raw" if world == 1 else "cutesdr_bank_process_async_bcast (NCCL inside libcutesdr_cuda)",
                        "mgpu": mg_info},
                "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu}
        if alt_res is not None:
            ar = alt_res["roof"] or {}
            line["alt_ingest"] = {
                "ingest": alt_res["ingest"], "value": alt_res["value"], "unit": "Msps*ch", "ms_per_block": alt_res["ms_max"] / args.steps / nblk,
                "e2e": {"value": alt_res["e2e_value"], "unit": "Msps*ch", "h2d_bytes_per_step": alt_res["h2d"], "d2h_bytes_per_step": alt_res["d2h"]},
                "roofline": {k: ar.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "launch_ms", "peak_source", "executed")},
                "note": "the same workload and run with the stream in the radio's other sample format (int16 I/Q words take kernel 1T's fp16 "
                        "form and halve the host->device bytes; complex64 takes the tf32 form)"}
        emit(line)
    del mg
    if world > 1:
        dist.destroy_process_group()
    return 0


def bank_stage_list(cs, w, mode, info):
    d = cs.CDownConvert()
    d.SetDataRate(w["in_rate"], M.max_bandwidth(mode, info))
    return d.stages()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS) + ["cfg2"])
    ap.add_argument("--ingest", default="cf32", choices=["cf32", "cs16"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--no-alt-ingest", action="store_true", help="skip the second measurement with the other sample format")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: everything native libraries print there (NCCL's banner, ...) is sent to
    # stderr for the duration of the run; emit() switches the real stdout back for the line itself
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.workload == "cfg2":
        if int(os.environ.get("RANK", "0")) != 0:
            return 0
        return run_cfg2(args)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


_REAL_STDOUT = None


def emit(line):
    sys.stdout.flush()
    if _REAL_STDOUT is not None:
        os.dup2(_REAL_STDOUT, 1)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    sys.exit(main())
