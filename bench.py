#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the fused FFT+waterfall path (BASELINE.json's metric).

One "step" = one pass of the hot path over one recording segment of BASELINE config 1 shape
(HackRF int8 IQ, 2^24 samples, 4096-point FFT, dB, avg = 8, peak hold, every waterfall row
stored) per GPU.  With N GPUs the long recording is time-sharded (config 5): rank r transforms
its own 2^24-sample segment, then the peak-hold and averaged spectra are reduced over NCCL
(weak scaling, no data-path collective, ONE packed all-reduce per recording).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`    : samples of ALL ranks / device time (CUDA events, max over ranks), inputs in HBM.
`e2e`      : same metric through the public API with pinned HOST buffers: H2D of the IQ bytes and
             D2H of rows/peaks/avg inside the timed region.
`roofline` : algorithmic HBM bytes of the fused kernel / its mean launch duration over
             >= 200 back-to-back launches, against the measured copy bandwidth in MEASURED_PEAKS.json.
`verified` : the run checks its OWN output after the timed region: sampled rows of the last step against
             the oracle, the reduced peak hold against the maximum over every rank's rows, the reduced
             average against the newest rows of the last rank.  A run that fails the check exits 1.
`configs`  : device-timed lines for the other BASELINE configs (C3 size sweep, C2 wFM, C4 nFM/USB/LSB/CW)
             with their algorithmic GB/s and the CPU arm beside them (rank 0, N = 1 only).
`recording`: BASELINE config 5 as written (strong scaling): ONE recording of 2^34 samples (or the largest
             power of two that fits the GPUs' memory) split over the ranks, time to the reduced result.
`cpu_baseline` / `--impl reference`: the reference's CPU path (its own pffft.c compiled in place
             + restated JVM stages, oracle/_ref) on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_FFT = 4096
SAMPLES = 1 << 24          # per GPU per step (BASELINE config 1)
FRAMES = SAMPLES // N_FFT
AVG_LEN = 8
FMT_S8 = 0
METRIC = "IQ Msamples/s through fused FFT+waterfall"
UNIT = "Msamples/s"
NBUF = 8                   # distinct input/output sets rotated through: 8 x 96 MiB >> 126 MB L2
DB_TOL = 0.01              # BASELINE.json: 0.01 dB on log spectra
# FP32-pipe work of the default N = 4096 kernel per warp and frame, from its SASS (profiles/r02_sass_hist_4096.txt):
# packed FADD2/FMUL2/FFMA2 occupy the pipe for two cycles, scalar FP32 for one
FP32_PIPE_CYCLES_PER_WARP_FRAME = 2 * 312 + 48


def workload_config(n_gpus, window):
    return {
        "workload": "BASELINE config 1 per GPU: HackRF int8 IQ, 2^24 samples @20 Msps -> %s 4096-pt FFT, dB, "
                    "avg=8, peak hold, all 4096 waterfall rows stored" % window
                    + ("" if n_gpus == 1 else "; config 5 style time-sharding: one 2^24-sample segment per GPU per "
                                              "step, peak-hold/average spectra reduced over NCCL once at the end of the timed recording"),
        "fft_size": N_FFT, "samples_per_gpu_per_step": SAMPLES, "frames_per_gpu_per_step": FRAMES,
        "avg_len": AVG_LEN, "peak_hold": True, "format": "int8 IQ", "window": window,
        "l2_policy": "rotating %d distinct input/output buffer sets (%.0f MiB) larger than the 126 MB L2"
                     % (NBUF, NBUF * (SAMPLES * 6) / 2 ** 20),
    }


class ClockSampler(threading.Thread):
    """nvidia-smi clock / throttle-reason samples while the GPU is under load."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None
        self.t0 = self.t1 = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.samples.append((time.time(), [x.strip() for x in line.split(",")]))
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sel = [s for t, s in self.samples if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e30) and len(s) >= 7]
        if not sel:
            sel = [s for _, s in self.samples if len(s) >= 7]
        if not sel:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[3 + i].lower() == "active" for s in sel)]
        mhz = [float(s[0]) for s in sel if s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in sel if s[1].replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(mhz) if mhz else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sel)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """Steady-state dram__bytes_read+write per launch of the fused kernel: an ncu RANGE capture over consecutive
    launches on rotating buffers (tools/traffic_capture.py; the write-back of one launch's rows happens inside
    the window of the next ones), committed under profiles/.  ncu cannot run inside this process, so the figure
    is the committed capture's; its file is named beside it."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return d.get("spectrum_kernel_dram_bytes_per_launch"), d.get("source")
        except Exception:
            return None, None
    return None, None


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pass(iq, nthreads, passes, n_fft=N_FFT, fmt=FMT_S8):
    """Msamples/s of the reference CPU path (oracle/_ref: reference pffft.c + nativedsp.cpp loop
    + restated JVM stages) over `passes` passes of the recording in `iq`."""
    import numpy as np
    from oracle import oracle as O
    samples = len(iq) // O.BYTES_PER_SAMPLE[fmt]
    frames = samples // n_fft
    rows = np.empty((frames, n_fft), np.float32)
    peaks = np.empty(n_fft, np.float32)
    avg = np.empty(n_fft, np.float32)
    R = O.ref()
    t0 = time.perf_counter()
    for _ in range(passes):
        R.ref_spectrum_run(fmt, iq, samples, n_fft, AVG_LEN, rows.ctypes.data, peaks.ctypes.data,
                           avg.ctypes.data, nthreads)
    dt = time.perf_counter() - t0
    return samples * passes / dt / 1e6, dt


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np
    from oracle import oracle as O
    O.build()
    kind = "reference" if O.ref_available() else "port"
    iq = O.synth_iq(FMT_S8, SAMPLES)
    cores = host_cores()
    if kind == "reference":
        step = lambda: cpu_reference_pass(iq, cores, 1)
    else:
        def step():
            t0 = time.perf_counter()
            O.spectrum_run(FMT_S8, iq, N_FFT, AVG_LEN)
            dt = time.perf_counter() - t0
            return SAMPLES / dt / 1e6, dt
        cores = 1
    for _ in range(max(args.warmup, 1)):
        step()
    t = 0.0
    for _ in range(args.steps):
        t += step()[1]
    value = SAMPLES * args.steps / t / 1e6
    sample = "one pass over the full 2^24-sample recording per step, all %d host threads" % cores
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(1, "Blackman (reference)"),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))
    return 0


# --------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa_node(index):
    """Run this rank's host side (and first-touch its pinned buffers) on the CPUs next to its GPU: with eight ranks
    streaming 96 MiB per step each over PCIe, buffers on the far socket halve the end-to-end rate."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass  # no NVML or no affinity information: keep the inherited CPU set
    return 0


def sequential_mean(rows_newest_first, np):
    """AnalyzerSurface.kt:710-714: float32 sum newest -> oldest, one division."""
    s = np.zeros(rows_newest_first.shape[1], np.float32)
    for r in rows_newest_first:
        s = (s + r).astype(np.float32)
    return (s / np.float32(len(rows_newest_first))).astype(np.float32)


def verify_run(torch, dist, rfa, rank, world, rows, last_set, peaks_out, avg_out, first_sample_of):
    """Checks the timed region's results (copies taken right after it).  Returns (ok, details)."""
    import numpy as np
    detail = {}
    ok = True
    # (1) sampled rows of the last step against the oracle, regenerated on the CPU from the sample index
    from oracle import oracle as O
    last_rows = rows[last_set]
    sample_frames = [0, 1, 2047, FRAMES - 2, FRAMES - 1]
    worst = 0.0
    for f in sample_frames:
        iq = O.synth_iq(FMT_S8, N_FFT, first=first_sample_of(last_set) + f * N_FFT)
        want, _, _ = O.spectrum_run(FMT_S8, iq, N_FFT, 0)
        got = last_rows[f].cpu().numpy()
        worst = max(worst, float(np.abs(got - want[0]).max()))
    detail["rows_max_abs_db_error_vs_oracle"] = worst
    detail["rows_frames_checked"] = len(sample_frames)
    rows_ok = worst < DB_TOL
    # (2) peak hold = element-wise maximum over EVERY row this rank produced (all NBUF sets were transformed), then over ranks
    local_max = torch.stack([r.max(dim=0).values for r in rows]).max(dim=0).values
    # (3) averaged spectrum = the newest AVG_LEN+1 rows of the LAST rank's last segment
    tail = torch.flip(last_rows[FRAMES - (AVG_LEN + 1):], dims=[0]).contiguous()
    if world > 1:
        gathered = [torch.empty_like(local_max) for _ in range(world)] if rank == 0 else None
        dist.gather(local_max, gathered, dst=0)
        dist.broadcast(tail, src=world - 1)
        if rank == 0:
            local_max = torch.stack(gathered).max(dim=0).values
    peaks_ok = avg_ok = True
    if rank == 0:
        peaks_ok = bool(torch.equal(peaks_out, local_max))
        want_avg = sequential_mean(tail.cpu().numpy(), np)
        avg_ok = bool(np.array_equal(avg_out.cpu().numpy(), want_avg))
        detail["peaks_equal_max_over_all_ranks_rows"] = peaks_ok
        detail["avg_equals_mean_of_newest_rows_of_last_rank"] = avg_ok
    flags = torch.tensor([1.0 if (rows_ok and peaks_ok and avg_ok) else 0.0], device="cuda")
    if world > 1:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    ok = bool(flags.item() == 1.0)
    detail["rows_within_0.01_db_on_every_rank"] = ok or (rows_ok and world == 1)
    return ok, detail


def time_calls(torch, stream, fn, reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for k in range(reps):
        fn(k)
    b.record(stream)
    stream.synchronize()
    return a.elapsed_time(b) / reps


def other_configs(torch, rfa, ctx, stream, peak_gbs, with_cpu):
    """Device-timed throughput of BASELINE configs 3, 2 and 4 (one GPU), the CPU arm beside each."""
    import numpy as np
    out = {}
    S = SAMPLES
    cores = host_cores()
    O = None
    if with_cpu:
        from oracle import oracle as O_
        O = O_
    # ---- C3: FFT-size sweep on 20 Msps int8 IQ, rows + peak hold + average, 2^24 samples per launch
    nset = 6
    with torch.cuda.stream(stream):
        iqs = [torch.empty(S * 2, dtype=torch.uint8, device="cuda") for _ in range(nset)]
        for j, b in enumerate(iqs):
            rfa.synth_iq(ctx, rfa.FMT_S8, S, b, first=j * S)
        rows = [torch.empty(S, dtype=torch.float32, device="cuda") for _ in range(nset)]
        # (n, cluster): the last two entries time the thread-block-cluster path for N >= 32768 (knob "cluster" = 1: the
        # four-step intermediate in distributed shared memory -- a sixth of the DRAM traffic, slower)
        for n, cluster in ((1024, 0), (2048, 0), (8192, 0), (16384, 0), (32768, 0), (65536, 0), (32768, 1), (65536, 1)):
            frames = S // n
            ctx.set_option("cluster", cluster)
            plan = rfa.SpectrumPlan(ctx, rfa.FMT_S8, n, avg_len=AVG_LEN, peak_hold=True)
            peaks = torch.zeros(n, dtype=torch.float32, device="cuda")
            avg = torch.zeros(n, dtype=torch.float32, device="cuda")
            views = [r.view(frames, n) for r in rows]

            def call(k, plan=plan, views=views, peaks=peaks, avg=avg, frames=frames):
                plan.process(iqs[k % nset], frames, rows=views[k % nset], peaks=peaks, avg=avg, peaks_accumulate=True)
            for k in range(nset):
                call(k)
            stream.synchronize()
            l0 = ctx.launch_count
            ms = time_calls(torch, stream, call, 60)
            alg = plan.algorithmic_bytes(frames, True)
            line = {"workload": "C3: int8 IQ @20 Msps, %d-pt FFT, rows + peak hold + avg=8, 2^24 samples per call%s" % (
                        n, " (cluster path: intermediate in distributed shared memory)" if cluster else ""),
                    "Msamples_per_s": S / ms / 1e3, "us_per_call": ms * 1e3, "algorithmic_GBps": alg / ms / 1e6,
                    "frac_of_hbm_peak": alg / ms / 1e6 / peak_gbs, "launches_per_call": (ctx.launch_count - l0) / 60}
            if O is not None and O.ref_available() and not cluster:
                h = O.synth_iq(rfa.FMT_S8, 1 << 22)
                cpu_reference_pass(h, cores, 1, n)
                v, dt = cpu_reference_pass(h, cores, 4, n)
                line["cpu_reference"] = {"Msamples_per_s": v, "cores": cores, "sample": "4 passes over 2^22 samples (%.2f s)" % dt}
            out["C3_fft_%d%s" % (n, "_cluster" if cluster else "")] = line
            plan.close()
        ctx.set_option("cluster", 0)
        del rows, iqs
    # ---- C2 / C4: IQ -> audio chains, 2^24 samples per call, device-resident
    chains = [("C2_wfm", "C2: RTL-SDR uint8 @2.4 Msps -> mixer + resampler to 384 kHz -> wFM -> 48 kHz audio", rfa.FMT_U8, 2_400_000, rfa.MODE_WFM, 100_000, 8192),
              ("C4_nfm", "C4: Airspy int16 @10 Msps -> channel extraction -> nFM", rfa.FMT_S16LE, 10_000_000, rfa.MODE_NFM, 10_000, 65536),
              ("C4_usb", "C4: Airspy int16 @10 Msps -> channel extraction -> USB", rfa.FMT_S16LE, 10_000_000, rfa.MODE_USB, 2_800, 65536),
              ("C4_lsb", "C4: Airspy int16 @10 Msps -> channel extraction -> LSB", rfa.FMT_S16LE, 10_000_000, rfa.MODE_LSB, 2_800, 65536),
              ("C4_cw", "C4: Airspy int16 @10 Msps -> channel extraction -> CW", rfa.FMT_S16LE, 10_000_000, rfa.MODE_CW, 300, 65536)]
    for key, name, fmt, fs, mode, width, packet in chains:
        off = fs // 10
        mul = 256 if fmt == rfa.FMT_S16LE else 1
        comps = [(rfa.synth_step(off / fs), 60 * mul, rfa.synth_step(1000 / fs), 3_130_000 if mode in (2, 3) else 0),
                 (rfa.synth_step((off + 1200) / fs), 20 * mul, 0, 0)]
        bps = rfa.BYTES_PER_SAMPLE[fmt]
        with torch.cuda.stream(stream):
            iq = torch.empty(S * bps, dtype=torch.uint8, device="cuda")
            rfa.synth_iq(ctx, fmt, S, iq, comps=comps, noise_shift=3)
            plan = rfa.ChainPlan(ctx, fmt, fs, 100_000_000, 100_000_000 + off, mode, width, packet, 1.0, rfa.SUM_FMA)
            audio = torch.empty(plan.max_audio(S), dtype=torch.float32, device="cuda")
            n_audio = [0]

            def call(k, plan=plan, iq=iq, audio=audio, n_audio=n_audio):
                n_audio[0] = plan.process(iq, S, audio)
            call(0)
            stream.synchronize()
            l0 = ctx.launch_count
            ms = time_calls(torch, stream, call, 10)
            alg = S * bps + 4 * n_audio[0]
            line = {"workload": name + ", 2^24 samples per call", "Msamples_per_s": S / ms / 1e3, "us_per_call": ms * 1e3,
                    "algorithmic_GBps": alg / ms / 1e6, "frac_of_hbm_peak": alg / ms / 1e6 / peak_gbs,
                    "launches_per_call": (ctx.launch_count - l0) / 10,
                    "resampler": "I/D = %d/%d, %d taps per phase" % (plan.interpolation, plan.decimation, plan.taps_per_phase)}
            plan.close()
            del iq, audio
        if O is not None:
            ns = 1 << 21
            h = O.synth_iq(fmt, ns, comps=comps, noise_shift=3)
            t0 = time.perf_counter()
            O.chain_run(fmt, h, fs, 100_000_000, 100_000_000 + off, mode, width, packet)
            dt = time.perf_counter() - t0
            line["cpu_port"] = {"Msamples_per_s": ns / dt / 1e6, "cores": 1, "sample": "2^21 samples on one thread (%.2f s)" % dt}
        out[key] = line
    return out


def recording_strong_scaling(torch, dist, rfa, ctx, stream, plan, shard_cls, rank, world, log2_total):
    """BASELINE config 5 as written: one recording of 2^log2_total samples, rank r holds and transforms samples
    [r * 2^log2/world, (r+1) * 2^log2/world) in 2^24-sample calls, every waterfall row of the recording is stored,
    peak hold and average are exchanged in one collective at the end.  Returns (seconds, log2 actually run)."""
    free, _ = torch.cuda.mem_get_info()
    # per sample: 2 B of IQ + 4 B of rows resident on the owning GPU
    while log2_total > 26 and (1 << log2_total) // world * 6 > free - (6 << 30):
        log2_total -= 1
    if world > 1:  # every rank must run the same recording
        t = torch.tensor([log2_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        log2_total = int(t.item())
    per_rank = (1 << log2_total) // world
    calls = per_rank // SAMPLES
    shard = shard_cls(plan, rank, world)
    with torch.cuda.stream(stream):
        iq = torch.empty(per_rank * 2, dtype=torch.uint8, device="cuda")
        rows = torch.empty((per_rank // N_FFT, N_FFT), dtype=torch.float32, device="cuda")
        for c in range(calls):
            rfa.synth_iq(ctx, rfa.FMT_S8, SAMPLES, iq[c * SAMPLES * 2:(c + 1) * SAMPLES * 2], first=rank * per_rank + c * SAMPLES)
        peaks = torch.full((N_FFT,), -999999.0, dtype=torch.float32, device="cuda")
        avg = torch.zeros(N_FFT, dtype=torch.float32, device="cuda")
        total_frames = (1 << log2_total) // N_FFT

        def run():
            for c in range(calls):
                plan.process(iq[c * SAMPLES * 2:(c + 1) * SAMPLES * 2], FRAMES, rows=rows[c * FRAMES:(c + 1) * FRAMES],
                             peaks=peaks, avg=avg, peaks_accumulate=c > 0)
            shard.reduce(total_frames, rows[(calls - 1) * FRAMES:], peaks, avg)
        run()   # warm: pages touched, NCCL connected
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        run()
        e1.record(stream)
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        # the result of the sharded pass: the newest rows live on the last rank
        ok = True
        if rank == world - 1:
            import numpy as np
            tail = torch.flip(rows[-(AVG_LEN + 1):], dims=[0]).cpu().numpy()
            ok = bool(np.array_equal(avg.cpu().numpy(), sequential_mean(tail, np)))
        del iq, rows
    return float(ms.item()) * 1e-3, log2_total, calls, ok


def chain_strong_scaling(torch, dist, rfa, ctx, stream, rank, world, log2_total=29):
    """BASELINE config 4 over the box: ONE Airspy int16 recording of 2^log2_total samples @10 Msps, time-sharded by
    whole packets with a warm-up halo (rfanalyzer_b200.sharding.ShardedChain: seek to a packet boundary, re-process the
    halo, discard its audio; no data-path collective), nFM and USB.  The run checks itself: every rank's audio is
    gathered on rank 0 and compared with the sequential single-GPU run of the whole recording computed there, to
    BASELINE's tolerance for audio (1e-4 of the peak).  The delay lines are exact after the halo, but the stripe
    resampler's summation order per output depends on where its tiles fall in the call (which outputs share a warp
    task, how a task's sample range is split), so a rank's bits equal the sequential run's only when its segment starts
    on a tile boundary; the AGC maximum of USB decays by 0.95 per packet over the 256-packet halo."""
    from rfanalyzer_b200.sharding import ShardedChain
    fs, packet, total = 10_000_000, 65536, 1 << log2_total
    off = fs // 10
    out = {}
    for name, mode, width in (("nfm", rfa.MODE_NFM, 10_000), ("usb", rfa.MODE_USB, 2_800)):
        comps = [(rfa.synth_step(off / fs), 60 * 256, rfa.synth_step(1000 / fs), 3_130_000 if mode == rfa.MODE_NFM else 0),
                 (rfa.synth_step((off + 1200) / fs), 20 * 256, 0, 0)]
        with torch.cuda.stream(stream):
            plan = rfa.ChainPlan(ctx, rfa.FMT_S16LE, fs, 100_000_000, 100_000_000 + off, mode, width, packet, 1.0, rfa.SUM_FMA)
            sc = ShardedChain(plan, rank, world)
            halo_start, first, n = sc.segment(total)
            mine = first + n - halo_start
            iq = torch.empty(mine * 4, dtype=torch.uint8, device="cuda")
            rfa.synth_iq(ctx, rfa.FMT_S16LE, mine, iq, first=halo_start, comps=comps, noise_shift=3)
            audio = torch.zeros(plan.max_audio(max(n, first - halo_start, 1)), dtype=torch.float32, device="cuda")
            sc.process(iq, total, audio)  # warm
            stream.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            index, got = sc.process(iq, total, audio)
            e1.record(stream)
            torch.cuda.synchronize()
            ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
            meta = torch.tensor([index, got], dtype=torch.int64, device="cuda")
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                metas = [torch.zeros_like(meta) for _ in range(world)]
                dist.all_gather(metas, meta)
                cap = int(max(int(m[1]) for m in metas))
                padded = torch.zeros(cap, dtype=torch.float32, device="cuda")
                padded[:got] = audio[:got]
                pieces = [torch.zeros_like(padded) for _ in range(world)] if rank == 0 else None
                dist.gather(padded, pieces, dst=0)
            else:
                metas, pieces = [meta], [audio[:got].clone()]
            ok, detail = True, None
            if rank == 0:
                del iq
                whole = torch.empty(total * 4, dtype=torch.uint8, device="cuda")
                rfa.synth_iq(ctx, rfa.FMT_S16LE, total, whole, comps=comps, noise_shift=3)
                seq = rfa.ChainPlan(ctx, rfa.FMT_S16LE, fs, 100_000_000, 100_000_000 + off, mode, width, packet, 1.0, rfa.SUM_FMA)
                want = torch.zeros(seq.max_audio(total), dtype=torch.float32, device="cuda")
                nw = seq.process(whole, total, want)
                stream.synchronize()
                pos = 0
                worst = 0.0
                identical = True
                peak = float(want[:nw].abs().max().item())
                for r in range(world):
                    idx, cnt = int(metas[r][0]), int(metas[r][1])
                    ok = ok and idx == pos
                    a, b = pieces[r][:cnt], want[idx:idx + cnt]
                    identical = identical and bool(torch.equal(a, b))
                    if cnt:
                        worst = max(worst, float((a - b).abs().max().item()))
                    pos += cnt
                ok = ok and pos == nw and worst <= 1e-4 * peak
                detail = "max |diff| %.2e of peak%s" % (worst / peak if peak else 0.0, ", bit-identical" if identical else "")
                seq.close()
                del whole, want
            plan.close()
            secs = float(ms.item()) * 1e-3
            out[name] = {"Msamples_per_s": total / secs / 1e6, "seconds": secs, "halo_packets": sc.halo_packets,
                         "audio_samples": int(sum(int(m[1]) for m in metas)), "equals_sequential_run": bool(ok), "check": detail}
        torch.cuda.empty_cache()
    return {"workload": "BASELINE config 4 over the box: ONE Airspy int16 recording of 2^%d samples @10 Msps time-sharded over %d "
                        "GPU(s) by whole packets with a warm-up halo, no data-path collective; device-timed, max over ranks"
                        % (log2_total, world), "n_gpus": world, "scaling": "strong", **out}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--window", default="blackman", choices=["blackman", "hann"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--recording-log2", type=int, default=34, help="config 5: log2 of the recording's samples (0 = skip)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the C2/C3/C4 lines")
    ap.add_argument("--no-clock-probe", action="store_true", help="skip the 1 s clock-sampling loop (ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)
    # stdout carries exactly ONE JSON line: libraries that print to file descriptor 1 (NCCL's version banner does,
    # whatever NCCL_DEBUG_FILE says) are sent to stderr for the whole run, the JSON line goes to the saved descriptor
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)

    import numpy as np
    import torch
    import torch.distributed as dist
    import rfanalyzer_b200 as rfa
    from rfanalyzer_b200.sharding import ShardedSpectrum

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("WORLD_SIZE %d does not match --gpus %d" % (world, args.gpus))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch with: python -m torch.distributed.run --nnodes=1 --nproc-per-node %d "
                         "--master-addr 127.0.0.1 --master-port P bench.py --gpus %d ..." % (args.gpus, args.gpus))
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # stdout carries exactly one JSON line: whatever NCCL logs (NCCL_DEBUG=VERSION/INFO on some boxes) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    window = rfa.WIN_HANN if args.window == "hann" else rfa.WIN_BLACKMAN_REF
    stream = torch.cuda.Stream()
    ctx = rfa.Context(local_rank, stream)
    plan = rfa.SpectrumPlan(ctx, rfa.FMT_S8, N_FFT, window=window, avg_len=AVG_LEN, peak_hold=True)
    shard = ShardedSpectrum(plan, rank, world)
    total_frames = FRAMES * world

    def barrier():
        if world > 1:
            dist.barrier()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()

    def first_sample_of(j):
        # the long recording: buffer set j, rank r owns samples [(j*world + r) * 2^24, +2^24)
        return (j * world + rank) * SAMPLES

    with torch.cuda.stream(stream):
        iqs, rows = [], []
        for j in range(NBUF):
            buf = torch.empty(SAMPLES * 2, dtype=torch.uint8, device="cuda")
            rfa.synth_iq(ctx, rfa.FMT_S8, SAMPLES, buf, first=first_sample_of(j))
            iqs.append(buf)
            rows.append(torch.empty((FRAMES, N_FFT), dtype=torch.float32, device="cuda"))
        peaks = torch.full((N_FFT,), -999999.0, dtype=torch.float32, device="cuda")
        avg = torch.zeros(N_FFT, dtype=torch.float32, device="cuda")

        def step(k):
            # each rank transforms its segment; peak hold / average accumulate locally and are
            # exchanged once, when the recording (= the timed region) ends -- config 5's reduction
            shard.process(iqs[k % NBUF], total_frames, rows[k % NBUF], peaks, avg, peaks_accumulate=True, reduce=False)

        for k in range(max(args.warmup, NBUF)):   # every buffer set is transformed at least once (the peak check relies on it)
            step(k)
        shard.reduce(total_frames, rows[0], peaks, avg)  # warm-up of the collective too (NCCL connects lazily)
        stream.synchronize()
        # ~1 s of the same steps so that nvidia-smi (100 ms period) samples clocks UNDER THIS LOAD
        if sampler:
            sampler.t0 = time.time()
        t_end = time.time() + (0.0 if args.no_clock_probe else 1.0)
        k = 0
        while time.time() < t_end:
            for _ in range(50):
                step(k)
                k += 1
            stream.synchronize()

        # ---- timed region: device time, inputs resident in HBM --------------------------------
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        launches0 = ctx.launch_count
        ev0.record(stream)
        for k in range(args.steps):
            step(k)
        last_set = (args.steps - 1) % NBUF
        shard.reduce(total_frames, rows[last_set], peaks, avg)  # NCCL, inside the timed region
        ev1.record(stream)
        torch.cuda.synchronize()
        barrier()
        if sampler:
            sampler.t1 = time.time()
        launches = ctx.launch_count - launches0
        ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        total_ms = float(ms.item())
        peaks_out, avg_out = peaks.clone(), avg.clone()   # the timed region's results, checked below

        # ---- the run checks its own output ---------------------------------------------------------
        verified, verify_detail = verify_run(torch, dist, rfa, rank, world, rows, last_set, peaks_out, avg_out, first_sample_of)

        # ---- kernel-only duration (roofline numerator): one event pair around >= 200 back-to-back launches of
        # the fused kernel on its own stream (no collective, no host work in between), whatever --steps is -----
        nk = max(args.steps, 200)
        kernel_ms = time_calls(torch, stream, lambda k: plan.process(iqs[k % NBUF], FRAMES, rows=rows[k % NBUF], peaks=peaks,
                                                                     avg=avg, peaks_accumulate=True), nk)

    # ---- end to end: pinned host buffers through the public API ---------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    h_iq = [torch.empty(SAMPLES * 2, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    h_rows = [torch.empty((FRAMES, N_FFT), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    h_peaks = torch.empty(N_FFT, dtype=torch.float32, pin_memory=True)
    h_avg = torch.empty(N_FFT, dtype=torch.float32, pin_memory=True)
    for j in range(2):
        h_iq[j].copy_(iqs[j])
    torch.cuda.synchronize()
    d_sum = torch.empty(2 * N_FFT, dtype=torch.float32, device="cuda")

    def e2e_step(k):
        plan.process(h_iq[k % 2], FRAMES, rows=h_rows[k % 2], peaks=h_peaks, avg=h_avg, peaks_accumulate=k > 0)

    def e2e_exchange():
        # the summary exchange of the sharded pass, ONCE per recording like the device-timed loop: the host
        # results go up, one packed all-reduce, the reduced summaries come back
        if world > 1:
            with torch.cuda.stream(stream):
                d_sum[:N_FFT].copy_(h_peaks, non_blocking=True)
                if rank == world - 1:
                    d_sum[N_FFT:].copy_(h_avg, non_blocking=True)
                else:
                    d_sum[N_FFT:].fill_(float("-inf"))
                dist.all_reduce(d_sum, op=dist.ReduceOp.MAX)
                h_peaks.copy_(d_sum[:N_FFT], non_blocking=True)
                h_avg.copy_(d_sum[N_FFT:], non_blocking=True)
                stream.synchronize()

    for k in range(3):
        e2e_step(k)
    e2e_exchange()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    e2e_exchange()
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = SAMPLES * world * e2e_steps / float(e2e_s.item()) / 1e6
    # the host-buffer path must have produced the same rows as the device path (same kernels behind it)
    e2e_ok = bool(np.array_equal(h_rows[(e2e_steps - 1) % 2][:64].numpy(), rows[(e2e_steps - 1) % 2][:64].cpu().numpy()))
    # ---- the ceiling of that path on this box: the same bytes per step as bare pinned copies, no kernels ---------
    # (H2D of the IQ on one stream, D2H of the rows on another, every rank at once: what the host links of the box
    # sustain when all N GPUs stream together; e2e above cannot exceed it)
    s_up, s_down = torch.cuda.Stream(), torch.cuda.Stream()
    d_iq_c = torch.empty(SAMPLES * 2, dtype=torch.uint8, device="cuda")
    d_rows_c = torch.empty((FRAMES, N_FFT), dtype=torch.float32, device="cuda")

    def copy_step(k):
        with torch.cuda.stream(s_up):
            d_iq_c.copy_(h_iq[k % 2], non_blocking=True)
        with torch.cuda.stream(s_down):
            h_rows[k % 2].copy_(d_rows_c, non_blocking=True)
    for k in range(3):
        copy_step(k)
    torch.cuda.synchronize()
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        copy_step(k)
    torch.cuda.synchronize()
    copy_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(copy_s, op=dist.ReduceOp.MAX)
    copy_ceiling = SAMPLES * world * e2e_steps / float(copy_s.item()) / 1e6
    copy_gbs = (SAMPLES * 6) * world * e2e_steps / float(copy_s.item()) / 1e9
    del h_iq, h_rows, d_iq_c, d_rows_c

    # ---- config 5 as written: strong scaling of one long recording (after the headline buffers are freed) ------
    recording = None
    peak_gbs, peak_src = measured_peak()
    configs = None
    del iqs, rows, step
    torch.cuda.empty_cache()
    if args.recording_log2 > 0:
        try:
            secs, log2_run, calls, rec_ok = recording_strong_scaling(torch, dist, rfa, ctx, stream, plan, ShardedSpectrum,
                                                                      rank, world, args.recording_log2)
            recording = {"workload": "BASELINE config 5: ONE int8 IQ recording of 2^%d samples time-sharded over %d GPU(s), "
                                     "4096-pt FFT, every waterfall row stored, peak hold + avg=8 reduced in one NCCL all-reduce"
                                     % (log2_run, world),
                         "log2_samples": log2_run, "n_gpus": world, "scaling": "strong", "seconds": secs,
                         "Msamples_per_s": (1 << log2_run) / secs / 1e6, "calls_per_gpu": calls,
                         "avg_equals_mean_of_newest_rows": rec_ok,
                         "requested_log2_samples": args.recording_log2}
            verified = verified and (rec_ok if rank == world - 1 else True)
        except Exception as e:  # never lose the headline over the secondary measurement
            recording = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    chain_rec = None
    if not args.no_configs:
        try:
            chain_rec = chain_strong_scaling(torch, dist, rfa, ctx, stream, rank, world)
            if rank == 0:
                verified = verified and all(v["equals_sequential_run"] for v in chain_rec.values() if isinstance(v, dict))
        except Exception as e:
            chain_rec = {"error": repr(e)[:300]}
        torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_configs:
        try:
            configs = other_configs(torch, rfa, ctx, stream, peak_gbs, not args.no_cpu_baseline)
        except Exception as e:
            configs = {"error": repr(e)[:300]}

    if sampler:
        sampler.stop()

    rc = 0
    if rank == 0:
        alg_bytes = plan.algorithmic_bytes(FRAMES, True)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        clocks = sampler.summary() if sampler else None
        traffic, traffic_src = ncu_traffic()
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        fp32_floor_us = FRAMES * 8 * FP32_PIPE_CYCLES_PER_WARP_FRAME / (ctx.sm_count * 4) / sm_mhz
        out = {
            "metric": METRIC, "value": SAMPLES * world * args.steps / (total_ms * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, "Blackman (reference)" if window == rfa.WIN_BLACKMAN_REF else "Hann"),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": SAMPLES * 2,
                    "d2h_bytes_per_step": SAMPLES * 4 + 2 * N_FFT * 4, "steps": e2e_steps,
                    "rows_equal_device_path": e2e_ok, "numa_bound_cpus": numa_cpus,
                    "bare_copy_ceiling": {"value": copy_ceiling, "unit": UNIT, "host_link_GBps_all_gpus": copy_gbs,
                                          "what": "the same H2D + D2H bytes per step as plain pinned cudaMemcpyAsync on two streams per "
                                                  "GPU, all ranks at once, no kernels: the box's host-link limit for this path"}},
            "gpu_launches": int(launches) * world,
            "verified": bool(verified and e2e_ok), "verify": verify_detail,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": traffic, "traffic_source": traffic_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_us_per_launch": kernel_ms * 1e3,
                         "kernel_launches_timed": nk, "peak_source": peak_src,
                         "fp32_pipe_floor_us": fp32_floor_us, "frac_of_fp32_pipe_floor": fp32_floor_us / (kernel_ms * 1e3),
                         "note": "at 6 B/sample the 4096-pt FFT is not HBM-bound on CUDA cores: the FP32 pipe alone needs "
                                 "fp32_pipe_floor_us (packed FADD2/FMUL2/FFMA2 counted two cycles, SASS histogram under profiles/), "
                                 "i.e. the kernel cannot pass fp32 floor / HBM time of this roofline; DESIGN.md 4.1"},
        }
        if recording is not None:
            out["recording"] = recording
        if configs is not None:
            out["configs"] = configs
        if chain_rec is not None:
            out["chain_recording"] = chain_rec
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import oracle as O
                O.build()
                h = O.synth_iq(FMT_S8, SAMPLES)
                cores = host_cores()
                if O.ref_available():
                    one, t_one = cpu_reference_pass(h, 1, 3)
                    est = max(t_one / 3 / cores * 1.5, 1e-3)       # seconds per all-core pass, guessed
                    passes = max(4, min(400, int(10.0 / est)))     # about 10 s of CPU work
                    cpu_reference_pass(h, cores, 2)                # threads and pages warm
                    allc, t_all = cpu_reference_pass(h, cores, passes)
                    out["cpu_baseline"] = {"value": allc, "unit": UNIT, "cores": cores, "kind": "reference",
                                           "single_thread_value": one,
                                           "sample": "full 2^24-sample recording: 3 passes on 1 thread (%.1f s), "
                                                     "%d passes on all %d threads (%.1f s)" % (t_one, passes, cores, t_all)}
                else:
                    t0 = time.perf_counter()
                    O.spectrum_run(FMT_S8, h, N_FFT, AVG_LEN)
                    dt = time.perf_counter() - t0
                    out["cpu_baseline"] = {"value": SAMPLES / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                                           "sample": "one pass over the 2^24-sample recording (%.1f s)" % dt}
            except Exception as e:  # the baseline is a report, never a reason to lose the GPU number
                out["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "failed: %r" % e}
        os.write(json_fd, (json.dumps(out) + "\n").encode())
        if not out["verified"]:
            print("bench.py: the run's own output failed verification: %r" % (verify_detail,), file=sys.stderr)
            rc = 1
    if world > 1:
        dist.destroy_process_group()
    return rc


if __name__ == "__main__":
    sys.exit(main())
