"""rfa_iqconverter_* (csrc/iqconv.cu) against the reference's own iqconverter_int16.c compiled in place (oracle/_ref) and
the C restatement: bit-exact int16 IQ for ADC-like noise, tones, full-scale square waves (int16 wrap-around of the DC
blocker), silence and constant input (its dead zone), ragged call lengths, state carried across calls, host and device
buffers."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _kernels(oracle):
    ks = [("synthetic", oracle.synthetic_hb_kernel(47)), ("short", oracle.synthetic_hb_kernel(15))]
    k = oracle.airspy_hb_kernel()
    if k is not None:
        ks.insert(0, ("airspy", k))
    return ks


def _signals(n, rng):
    t = np.arange(n)
    adc = (np.clip(np.round(2100 + 350 * rng.standard_normal(n)), 0, 4095).astype(np.int64) - 2048) << 4   # a 12-bit ADC at a sane level, DC offset
    adc_full = (rng.integers(0, 4096, n) - 2048) << 4    # uniform over the whole range: w and y wrap around all the time
    tone = np.round(9000 * np.cos(2 * np.pi * 0.1234 * t) + 300 * rng.standard_normal(n) + 1500)
    sq = np.where(t % 7 < 3, -32768, 32767)
    burst = np.zeros(n)
    burst[n // 3: n // 3 + 5000] = rng.integers(-20000, 20000, 5000)      # silence, activity, silence again
    const = np.full(n, 1234)
    const[:100] = rng.integers(-500, 500, 100)
    return {"adc": adc, "adc_fullscale_noise": adc_full, "tone": tone, "square_fullscale": sq, "zero": np.zeros(n), "burst": burst, "const": const,
            "rand_fullscale": rng.integers(-32768, 32768, n)}


def _reference(oracle, kernel, x, cuts):
    use_ref = oracle.ref_available()
    cv = oracle.IqConverterInt16(kernel, use_ref=use_ref)
    out, a = [], 0
    for b in cuts + [len(x)]:
        out.append(cv.process(x[a:b].copy()))
        a = b
    return np.concatenate(out)


@pytest.mark.parametrize("device", [True, False])
def test_iqconverter_bit_exact(gpu_ctx, oracle, device):
    import torch
    import rfanalyzer_b200 as rfa
    n = 200_000
    rng = np.random.default_rng(7)
    cuts = [4, 5000, 5004, 131072]                      # ragged calls (all multiples of four)
    for kname, kernel in _kernels(oracle):
        for sname, sig in _signals(n, rng).items():
            x = sig.astype(np.int16)
            want = _reference(oracle, kernel, x, cuts)
            cv = rfa.IqConverterInt16(gpu_ctx, kernel)
            got, a = [], 0
            for b in cuts + [n]:
                part = x[a:b].copy()
                if device:
                    with torch.cuda.stream(gpu_ctx.torch_stream):
                        d = torch.from_numpy(part).cuda()
                        cv.process(d)
                        gpu_ctx.sync()
                        part = d.cpu().numpy()
                else:
                    cv.process(part)
                got.append(part)
                a = b
            got = np.concatenate(got)
            bad = np.nonzero(got != want)[0]
            assert bad.size == 0, (kname, sname, device, bad[:5], got[bad[:5]], want[bad[:5]])
            chunks, rerun, dead = cv.stats()
            assert chunks > 0
            if sname in ("adc", "tone"):
                assert rerun == 0 and dead == 0, (sname, rerun, dead)   # ordinary signals never reach the repair kernel


def test_port_equals_compiled_reference_on_the_gpu_inputs(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref is not built")
    rng = np.random.default_rng(3)
    k = oracle.airspy_hb_kernel()
    for sig in _signals(50_000, rng).values():
        x = sig.astype(np.int16)
        a = oracle.IqConverterInt16(k, use_ref=False).process(x.copy())
        b = oracle.IqConverterInt16(k, use_ref=True).process(x.copy())
        assert np.array_equal(a, b)


def test_airspy_raw_words_to_spectrum(gpu_ctx, oracle):
    """The path a raw Airspy recording takes: 12-bit ADC words -> convert_samples_int16 -> iqconverter -> the fused
    spectrum kernel as int16 IQ, all on the device; rows equal the oracle's rows of the reference converter's output."""
    import torch
    import rfanalyzer_b200 as rfa
    n_fft, frames = 4096, 12
    nreal = 2 * n_fft * frames
    rng = np.random.default_rng(11)
    t = np.arange(nreal)
    raw = np.clip(np.round(2048 + 900 * np.cos(2 * np.pi * 0.31 * t) + 40 * rng.standard_normal(nreal)), 0, 4095).astype(np.uint16)
    kernel = oracle.airspy_hb_kernel()
    if kernel is None:
        kernel = oracle.synthetic_hb_kernel(47)
    ref16 = np.empty(nreal, np.int16)
    oracle.lib().orc_airspy_convert_samples(raw.ctypes.data, ref16.ctypes.data, nreal)
    iq_ref = _reference(oracle, kernel, ref16, [])
    want, _, _ = oracle.spectrum_run(2, iq_ref.view(np.uint8), n_fft, 0)
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d_raw = torch.from_numpy(raw.view(np.int16)).cuda()
        d16 = torch.empty(nreal, dtype=torch.int16, device="cuda")
        rfa.airspy_convert_samples(gpu_ctx, d_raw, d16)
        rfa.IqConverterInt16(gpu_ctx, kernel).process(d16)
        plan = rfa.SpectrumPlan(gpu_ctx, rfa.FMT_S16LE, n_fft)
        rows = torch.zeros((frames, n_fft), dtype=torch.float32, device="cuda")
        plan.process(d16.view(torch.uint8), frames, rows=rows)
        gpu_ctx.sync()
    assert np.array_equal(d16.cpu().numpy(), iq_ref)
    assert np.abs(rows.cpu().numpy() - want).max() < 0.01


def test_iqconverter_throughput_smoke(gpu_ctx, oracle):
    """2^24 real samples in one call: a sanity bound on the speculative scheme (never the sequential rate)."""
    import time
    import torch
    import rfanalyzer_b200 as rfa
    n = 1 << 24
    with torch.cuda.stream(gpu_ctx.torch_stream):
        x = ((torch.randint(1500, 2700, (n,), device="cuda", dtype=torch.int32) - 2048) * 16).to(torch.int16)
        cv = rfa.IqConverterInt16(gpu_ctx, oracle.synthetic_hb_kernel(47))
        cv.process(x.clone())
        gpu_ctx.sync()
        y = x.clone()
        t0 = time.perf_counter()
        cv.process(y)
        gpu_ctx.sync()
        dt = time.perf_counter() - t0
    chunks, rerun, dead = cv.stats()
    print("iqconverter: 2^24 real samples in %.3f ms (%.1f Gsamples/s), chunks %d rerun %d" % (dt * 1e3, n / dt / 1e9, chunks, rerun))
    assert rerun == 0
    assert dt < 0.05
