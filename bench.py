#!/usr/bin/env python3
"""bench.py -- headline benchmark of the receive-chain bank: input Msps x channels per B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg5|cfg5full|cfg3]

One STEP = one DSP block (block_length wideband samples, ~10 ms of signal) pushed through the whole
chain -- NCO mix + CIC/half-band cascade, overlap-save FIR, S-meter, AGC, demodulator (+ resampler)
-- for every channel of the rank's bank. Default workload: BASELINE config 4, 1024-channel NBFM +
CFractResampler to 48 kHz on the "100 Msps" stream (100 147 200 sps), 1024 channels PER
GPU (weak scaling: every GPU sees the same wideband stream and owns its own 1024-channel slice).

  value   : device-timed, wideband blocks already resident in HBM (cycled through more distinct
            blocks than fit in L2).
  e2e     : the same through cutesdr_bank_process with HOST buffers: H2D of the block (rank 0, then an
            NCCL broadcast when N > 1) and D2H of every channel's audio inside the timed region.
  roofline: kernel 1. On the tensor-core path (k_mix_tc) bound = "tensor": the GEMM's algorithmic flops per launch /
            its CUDA-event time vs the measured tf32 peak (bf16 sustained / 2); the SURVEY 8(d) HBM streaming-model
            figure (8 B x samples x channels) rides along as roofline.hbm_model. CUDA-core path: bound = "hbm" (model).
  cpu_baseline / --impl reference : the UNMODIFIED reference dsp/*.cpp (oracle/_ref) on all host cores,
            one CDemodulator (+CFractResampler) per channel, on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cutesdr_b200 import modes as M  # noqa: E402

WORKLOADS = {
    # name: (in_rate, channels per GPU, mode picker, carrier spacing, audio_rate, description)
    "cfg4": (100147200.0, 1024, lambda c: M.DEMOD_FM, 78125.0, 48000.0,
             "cfg4: 1024-ch NBFM (+LP biquad, CFractResampler->48 kHz) on 100.1472 Msps, 1024 ch per GPU"),
    "cfg5": (200294400.0, 1024, lambda c: (M.DEMOD_AM, M.DEMOD_SAM, M.DEMOD_FM, M.DEMOD_USB)[c % 4], 39000.0, 0.0,
             "cfg5 slice: 1024-ch mixed AM/SAM/FM/USB with AGC on 200.2944 Msps, 1024 ch per GPU"),
    "cfg5full": (200294400.0, 4096, lambda c: (M.DEMOD_AM, M.DEMOD_SAM, M.DEMOD_FM, M.DEMOD_USB)[c % 4], 39000.0, 0.0,
                 "cfg5 at full width: 4096-ch mixed AM/SAM/FM/USB with AGC on 200.2944 Msps, 4096 ch per GPU"),
    "cfg3": (20000000.0, 256, lambda c: M.DEMOD_USB if c % 2 == 0 else M.DEMOD_LSB, 62500.0, 0.0,
             "cfg3: 256-ch USB/LSB SSB bank on 20 Msps, 256 ch per GPU"),
}


def channel_plan(name, rank, world):
    in_rate, nch, pick, spacing, audio_rate, desc = WORKLOADS[name]
    total = nch * world
    modes = [pick(c) for c in range(rank * nch, (rank + 1) * nch)]
    carriers = (np.arange(rank * nch, (rank + 1) * nch) - total / 2 + 0.5) * spacing
    infos = []
    for m in modes:
        if m == M.DEMOD_USB:
            infos.append(M.demod_info(m, HiCut=2800, LowCut=100))
        elif m == M.DEMOD_LSB:
            infos.append(M.demod_info(m, HiCut=-100, LowCut=-2800))
        else:
            infos.append(M.demod_info(m))
    return in_rate, nch, modes, carriers, infos, audio_rate, desc


def synth_blocks(L, nblocks, seed):
    """Cheap synthetic wideband stream: Gaussian noise floor + 64 AM/FM carriers, int16 full scale."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n = L * nblocks
    x = (300.0 * (rng.standard_normal(n, dtype=np.float32) + 1j * rng.standard_normal(n, dtype=np.float32))).astype(np.complex64)
    t = np.arange(n, dtype=np.float64)
    for k in range(8):
        f = (k - 3.5) * 0.031
        x += (1500.0 * (1.0 + 0.5 * np.cos(2 * np.pi * 2e-5 * (k + 1) * t)) * np.exp(2j * np.pi * f * t)).astype(np.complex64)
    return x.reshape(nblocks, L)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.idx = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([v.strip() for v in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_tf32_peak():
    """Dense tf32 tensor peak in TFLOP/s: half the bf16 figure (tcgen05 kind::tf32 runs at half the kind::f16 rate,
    B200_PROFILING.md). Kernel 1T is timed inside a long step, so the SUSTAINED bf16 number is the one halved."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d.get("bf16_tflops_sustained", d["bf16_tflops"])) / 2.0, "measured bf16 sustained / 2 (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 2250.0 / 2.0, "fallback: nominal 2.25 PFLOP/s bf16 / 2 (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(workload, steps, warmup, ch_per_core=4):
    """The reference's own CPU implementation on all host cores: one CDemodulator (+resampler) per
    channel, channels spread over std::threads; each step = one DSP block over a bounded channel count."""
    from oracle import ref_binding as rb
    in_rate, nch, modes, carriers, infos, audio_rate, desc = channel_plan(workload, 0, 1)
    big = in_rate > 30e6
    if not rb.ref_available(big=big):
        return None
    cores = os.cpu_count() or 1
    n_cpu_ch = min(nch, cores * ch_per_cpu_core(ch_per_core))
    sel = np.linspace(0, nch - 1, n_cpu_ch).astype(int)
    L = (int(in_rate / 100) & ~0xFF)
    blocks = synth_blocks(L, 1, seed=7)
    infos_by_mode = {}
    for c in sel:
        infos_by_mode[modes[c]] = infos[c]
    times = []
    for it in range(warmup + steps):
        # state is rebuilt per call (object construction is inside the timed call but negligible vs 1e6 samples)
        t, _ = rb.bench_chains([modes[c] for c in sel], [-carriers[c] for c in sel], infos_by_mode, in_rate, blocks[0], cores,
                               resample48k=audio_rate > 0, big=big)
        if it >= warmup:
            times.append(t)
    sec = float(np.sum(times))
    value = (L * n_cpu_ch * steps) / sec / 1e6
    return {"value": value, "cores": cores, "kind": "reference", "ms_per_step": 1e3 * sec / steps,
            "sample": "%d of %d channels x %d-sample block per step, %d steps, %d threads" % (n_cpu_ch, nch, L, steps, cores)}


def ch_per_cpu_core(default):
    return int(os.environ.get("CUTESDR_BENCH_CPU_CH_PER_CORE", default))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    steps = max(1, min(args.steps, 20))
    warm = min(args.warmup, 2)
    r = cpu_reference_run(args.workload, steps, warm)
    in_rate, nch, modes, carriers, infos, audio_rate, desc = channel_plan(args.workload, 0, 1)
    if r is None:
        # oracle/_ref did not travel: fall back to the C restatement on one core
        from oracle import oracle_binding as ob
        L = (int(in_rate / 100) & ~0xFF)
        blocks = synth_blocks(L, 1, seed=7)
        d = ob.Demodulator()
        d.SetInputSampleRate(in_rate)
        d.SetDemod(modes[0], infos[0])
        d.SetDemodFreq(-carriers[0])
        t0 = time.perf_counter()
        for _ in range(steps):
            d.run(blocks[0])
        sec = time.perf_counter() - t0
        r = {"value": L * steps / sec / 1e6, "cores": 1, "kind": "port", "ms_per_step": 1e3 * sec / steps,
             "sample": "1 channel x %d-sample block per step, %d steps, 1 thread (oracle port)" % (L, steps)}
    line = {"impl": "reference", "metric": "input_msps_x_channels", "value": r["value"], "unit": "Msps*ch", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "note": "reference CPU chain (dsp/*.cpp compiled headless), all host cores"},
            "cpu_baseline": {"value": r["value"], "unit": "Msps*ch", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
            "e2e": {"value": r["value"], "unit": "Msps*ch", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)
    return 0


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import cutesdr_b200 as cs

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    in_rate, nch, modes, carriers, infos, audio_rate, desc = channel_plan(args.workload, rank, world)
    bank = cs.ReceiverBank(nch, in_rate, device=local)
    if audio_rate > 0:
        bank.SetAudioRate(audio_rate)
    for c in range(nch):
        bank.SetDemod(c, modes[c], infos[c])
        bank.SetDemodFreq(c, -carriers[c])
    L = bank.block_length()
    stream = torch.cuda.ExternalStream(bank.stream(), device=dev)

    # distinct resident blocks: more than the 126 MB L2 can hold
    nblk = max(4, int(np.ceil(160e6 / (8.0 * L))))
    host_blocks = synth_blocks(L, nblk, seed=20260 + 4)
    h_pin = torch.from_numpy(host_blocks.view(np.float32).reshape(nblk, 2 * L)).pin_memory()
    d_blocks = h_pin.to(dev)
    audio_stride = 2304
    d_audio = torch.zeros((nch, audio_stride), dtype=torch.float32, device=dev)
    h_audio = torch.zeros((nch, audio_stride), dtype=torch.float32).pin_memory()
    n_out = np.zeros(nch, dtype=np.int32)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_device(i):
        b = d_blocks[i % nblk]
        return bank.process_device(b.data_ptr(), L, d_audio.data_ptr(), audio_stride)

    # ---- device-timed value
    for i in range(args.warmup):
        step_device(i)
    barrier()
    bank.kernel_timing(True)
    bank.kernel_time(0)
    launches0 = bank.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    produced = 0
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for i in range(args.steps):
            produced += step_device(args.warmup + i)
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
    value = (L * nch * world * args.steps) / (ms_max * 1e-3) / 1e6

    # ---- end to end: host block -> (rank 0 H2D -> NCCL broadcast) -> bank -> D2H audio
    e2e_steps = max(3, min(args.steps, 60))
    d_in = torch.empty(2 * L, dtype=torch.float32, device=dev)

    # two host audio landing buffers: a step's D2H may still be in flight when the next step is queued
    h_audio2 = [h_audio, torch.zeros((nch, audio_stride), dtype=torch.float32).pin_memory()]

    comm = torch.cuda.Stream(device=dev) if world > 1 else None

    def step_e2e(i):
        if world == 1:
            # pipelined C-ABI call: H2D of block i overlaps the kernels of block i-1, D2H on its own stream
            return bank.process_async_ptr(L, h_pin[i % nblk].data_ptr(), h_audio2[i & 1].data_ptr(), audio_stride, n_out)
        # N > 1: rank 0's H2D and the NCCL broadcast run on a communication stream; the bank takes the block over in
        # stream order (cutesdr_bank_process_async_device), so block i+1 is in flight while block i is processed and
        # the audio D2H runs on the bank's copy stream
        with torch.cuda.stream(comm):
            if rank == 0:
                d_in.copy_(h_pin[i % nblk], non_blocking=True)
            dist.broadcast(d_in, src=0)
            return bank.process_async_device_ptr(L, d_in.data_ptr(), comm.cuda_stream, h_audio2[i & 1].data_ptr(), audio_stride, n_out)

    for i in range(3):
        step_e2e(i)
    barrier()
    t0 = time.perf_counter()
    d2h = 0
    for i in range(e2e_steps):
        m = step_e2e(3 + i)
        d2h += int(m) * nch * 4
    bank.synchronize()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    te = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = (L * nch * world * e2e_steps) / float(te.item()) / 1e6

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        k1_bytes = 8.0 * L * nch           # per launch (one launch per chain group per block)
        roof = None
        if k1_n > 0 and k1_ms > 0:
            per_launch_s = (k1_ms / k1_n) * 1e-3
            groups = max(1, round(k1_n / max(1, args.steps)))
            ach = (k1_bytes / groups) / per_launch_s / 1e9
            traffic = None
            try:
                traffic = json.load(open(os.path.join(ROOT, "profiles", "k1_traffic.json")))[args.workload]["dram_bytes_per_launch"]
            except Exception:
                pass
            hbm_model = {"achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "peak_source": peak_src,
                         "note": "SURVEY 8(d) per-channel streaming model (8 B per sample*channel); NOT a DRAM figure -- every "
                                 "channel re-uses the staged samples, see traffic for the measured DRAM bytes"}
            on_tc, flops_block = bank.kernel_model(0)
            if on_tc and flops_block > 0:
                tpeak, tsrc = measured_tf32_peak()
                flops_launch = flops_block / groups
                ach_t = flops_launch / per_launch_s / 1e12
                roof = {"bound": "tensor", "achieved": ach_t, "peak": tpeak, "unit": "TFLOP/s", "frac": ach_t / tpeak, "traffic": traffic,
                        "kernel": "k_mix_tc", "launch_ms": k1_ms / k1_n, "launches": k1_n, "peak_source": tsrc,
                        "flops_per_launch": flops_launch,
                        "executed": {"tflops": 3.0 * ach_t, "frac": 3.0 * ach_t / tpeak,
                                     "note": "every fp32 product is three tf32 MMAs (hi*hi + hi*lo + lo*hi): the tensor pipe "
                                             "executes 3x the algorithmic flops"},
                        "hbm_model": hbm_model,
                        "note": "kernel 1T: NCO mix + 4 CIC3 stages as a complex GEMM [128 ch x 48 taps] x [48 x time] per channel "
                                "group; algorithmic flops = 2 x (4 x 48 real MACs) per channel and fs/16 output"}
            else:
                roof = dict(hbm_model)
                roof.update({"bound": "hbm", "traffic": traffic, "kernel": "k_mix_cic", "launch_ms": k1_ms / k1_n, "launches": k1_n,
                             "note": hbm_model["note"] + "; the CUDA-core kernel is FP32-issue bound"})
        cpu = None
        try:
            r = cpu_reference_run(args.workload, steps=3, warmup=1)
            if r:
                cpu = {"value": r["value"], "unit": "Msps*ch", "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        except Exception as e:  # the baseline is reported, never required
            cpu = {"value": None, "unit": "Msps*ch", "cores": os.cpu_count(), "kind": "reference", "sample": "failed: %r" % (e,)}
        line = {"metric": "input_msps_x_channels", "value": value, "unit": "Msps*ch", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "in_rate_sps": in_rate, "channels_per_gpu": nch, "block_length": L,
                           "l2": "cycles %d distinct %d-sample blocks (%.0f MB > 126 MB L2)" % (nblk, L, nblk * L * 8 / 1e6),
                           "realtime_factor": value / (in_rate * nch * world / 1e6),
                           "arithmetic": "float32 chain; kernel 1 on tensor cores = every fp32 product as 3 tf32 MMAs, fp32 accumulation"
                                         if (roof and roof.get("bound") == "tensor") else "float32 chain (CUDA cores)"},
                "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": "Msps*ch", "h2d_bytes_per_step": 8 * L, "d2h_bytes_per_step": d2h // e2e_steps,
                        "steps": e2e_steps},
                "gpu_launches": int(launches),
                "roofline": roof, "cpu_baseline": cpu}
        emit(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly one JSON line: everything native libraries print there (NCCL's version banner, ...) is
    # sent to stderr for the duration of the run; emit() switches the real stdout back for the line itself
    sys.stdout.flush()
    global _REAL_STDOUT
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
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
