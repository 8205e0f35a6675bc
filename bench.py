#!/usr/bin/env python
"""bench.py -- IQ Msamples/s through the fused FFT+waterfall path (BASELINE.json's metric).

One "step" = one pass of the hot path over one recording segment of BASELINE config 1 shape
(HackRF int8 IQ, 2^24 samples, 4096-point FFT, dB, avg = 8, peak hold, every waterfall row
stored) per GPU.  With N GPUs the long recording is time-sharded (config 5): rank r transforms
its own 2^24-sample segment, then the peak-hold and averaged spectra are reduced over NCCL
(weak scaling, no data-path collective).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`value`   : samples of ALL ranks / device time (CUDA events, max over ranks), inputs in HBM.
`e2e`     : same metric through the public API with pinned HOST buffers: H2D of the IQ bytes and
            D2H of rows/peaks/avg inside the timed region.
`roofline`: algorithmic HBM bytes of the fused kernel / its mean launch duration,
            against the measured copy bandwidth in MEASURED_PEAKS.json.
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
    """dram__bytes_read+write per launch of the fused kernel from the committed ncu capture."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)).get("spectrum_kernel_dram_bytes_per_launch")
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------- CPU reference arm
def cpu_reference_pass(iq, nthreads, passes):
    """Msamples/s of the reference CPU path (oracle/_ref: reference pffft.c + nativedsp.cpp loop
    + restated JVM stages) over `passes` passes of the 2^24-sample recording."""
    import numpy as np
    from oracle import oracle as O
    rows = np.empty((FRAMES, N_FFT), np.float32)
    peaks = np.empty(N_FFT, np.float32)
    avg = np.empty(N_FFT, np.float32)
    R = O.ref()
    t0 = time.perf_counter()
    for _ in range(passes):
        R.ref_spectrum_run(FMT_S8, iq, SAMPLES, N_FFT, AVG_LEN, rows.ctypes.data, peaks.ctypes.data,
                           avg.ctypes.data, nthreads)
    dt = time.perf_counter() - t0
    return SAMPLES * passes / dt / 1e6, dt


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
    except Exception:
        pass  # no NVML or no affinity information: keep the inherited CPU set


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--window", default="blackman", choices=["blackman", "hann"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clock-probe", action="store_true", help="skip the 1 s clock-sampling loop (ncu runs)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

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
    bind_to_gpu_numa_node(local_rank)
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

    with torch.cuda.stream(stream):
        # the long recording: step j, rank r owns samples [(j*world + r) * 2^24, +2^24)
        iqs, rows = [], []
        for j in range(NBUF):
            buf = torch.empty(SAMPLES * 2, dtype=torch.uint8, device="cuda")
            rfa.synth_iq(ctx, rfa.FMT_S8, SAMPLES, buf, first=(j * world + rank) * SAMPLES)
            iqs.append(buf)
            rows.append(torch.empty((FRAMES, N_FFT), dtype=torch.float32, device="cuda"))
        peaks = torch.full((N_FFT,), -999999.0, dtype=torch.float32, device="cuda")
        avg = torch.zeros(N_FFT, dtype=torch.float32, device="cuda")

        def step(k):
            # each rank transforms its segment; peak hold / average accumulate locally and are
            # exchanged once, when the recording (= the timed region) ends -- config 5's reduction
            shard.process(iqs[k % NBUF], total_frames, rows[k % NBUF], peaks, avg, peaks_accumulate=True, reduce=False)

        for k in range(args.warmup):
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
        shard.reduce(total_frames, rows[(args.steps - 1) % NBUF], peaks, avg)  # NCCL, inside the timed region
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

        # ---- kernel-only duration (roofline numerator): one event pair around a back-to-back batch of the
        # fused kernel's launches on its own stream (no collective, no host work in between) -------------
        nk = min(args.steps, 100)
        ka, kb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ka.record(stream)
        for k in range(nk):
            plan.process(iqs[k % NBUF], FRAMES, rows=rows[k % NBUF], peaks=peaks, avg=avg, peaks_accumulate=True)
        kb.record(stream)
        stream.synchronize()
        kernel_ms = ka.elapsed_time(kb) / nk

    # ---- end to end: pinned host buffers through the public API ---------------------------------
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    h_iq = [torch.empty(SAMPLES * 2, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    h_rows = [torch.empty((FRAMES, N_FFT), dtype=torch.float32, pin_memory=True) for _ in range(2)]
    h_peaks = torch.empty(N_FFT, dtype=torch.float32, pin_memory=True)
    h_avg = torch.empty(N_FFT, dtype=torch.float32, pin_memory=True)
    for j in range(2):
        h_iq[j].copy_(iqs[j])
    torch.cuda.synchronize()

    def e2e_step(k):
        plan.process(h_iq[k % 2], FRAMES, rows=h_rows[k % 2], peaks=h_peaks, avg=h_avg, peaks_accumulate=k > 0)
        if world > 1:  # the summary exchange of the sharded pass, from the host results
            with torch.cuda.stream(stream):
                dp = h_peaks.cuda(non_blocking=True)
                dist.all_reduce(dp, op=dist.ReduceOp.MAX)
                h_peaks.copy_(dp)
                stream.synchronize()

    for k in range(3):
        e2e_step(k)
    barrier()
    t0 = time.perf_counter()
    for k in range(e2e_steps):
        e2e_step(k)
    torch.cuda.synchronize()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = SAMPLES * world * e2e_steps / float(e2e_s.item()) / 1e6

    if sampler:
        sampler.stop()

    if rank == 0:
        peak_gbs, peak_src = measured_peak()
        alg_bytes = plan.algorithmic_bytes(FRAMES, True)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": SAMPLES * world * args.steps / (total_ms * 1e-3) / 1e6, "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world, "Blackman (reference)" if window == rfa.WIN_BLACKMAN_REF else "Hann"),
            "clocks": sampler.summary() if sampler else None,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": SAMPLES * 2,
                    "d2h_bytes_per_step": SAMPLES * 4 + 2 * N_FFT * 4, "steps": e2e_steps},
            "gpu_launches": int(launches) * world,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s",
                         "frac": achieved / peak_gbs, "traffic": ncu_traffic(),
                         "algorithmic_bytes_per_launch": alg_bytes, "kernel_us_per_launch": kernel_ms * 1e3,
                         "peak_source": peak_src,
                         "note": "at 6 B/sample the 4096-pt FFT is not HBM-bound on CUDA cores: issue slots cap it at ~0.55 of this "
                                 "roofline (packed FP32 instructions hold the scheduler two cycles), FP32 pipe 48% and shared-"
                                 "memory pipe 51% busy: DESIGN.md 4.1, profiles/r01b_*"},
        }
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
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
