"""K1/K2 on the GPU through the C ABI, bit-exact against the oracle (and therefore against the
reference's look-up tables): IQConverter.fillPacketIntoSamplePacket / mixPacketIntoSamplePacket."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def all_codes(fmt):
    if fmt == 2:  # every 16-bit code as I, reversed as Q
        i = np.arange(65536, dtype=np.uint16)
        q = i[::-1].copy()
        return np.stack([i, q], axis=1).reshape(-1).view(np.uint8).copy()
    i = np.arange(256, dtype=np.uint8)
    q = i[::-1].copy()
    return np.stack([i, q], axis=1).reshape(-1).copy()


def oracle_fill(oracle, fmt, iq, fs=2400000):
    L = oracle.lib()
    n = len(iq) // oracle.BYTES_PER_SAMPLE[fmt]
    c = L.orc_converter_new(fmt)
    L.orc_converter_set_sample_rate(c, fs)
    sp = oracle.PacketView(n)
    assert L.orc_converter_fill(c, iq, len(iq), sp.p) == n
    L.orc_converter_free(c)
    return sp.out_re(), sp.out_im()


@pytest.mark.parametrize("fmt", [0, 1, 2])
def test_convert_every_code_point_bit_exact(gpu_ctx, oracle, fmt):
    import torch
    iq = all_codes(fmt)
    n = len(iq) // oracle.BYTES_PER_SAMPLE[fmt]
    re_ref, im_ref = oracle_fill(oracle, fmt, iq)
    # host mode
    re, im = np.empty(n, np.float32), np.empty(n, np.float32)
    gpu_ctx.convert(fmt, iq, n, re, im)
    assert np.array_equal(re, re_ref) and np.array_equal(im, im_ref)
    # device mode
    with torch.cuda.stream(gpu_ctx.torch_stream):
        d = torch.from_numpy(iq).cuda()
        dre = torch.empty(n, dtype=torch.float32, device="cuda")
        dim = torch.empty(n, dtype=torch.float32, device="cuda")
        gpu_ctx.convert(fmt, d, n, dre, dim)
        gpu_ctx.sync()
    assert np.array_equal(dre.cpu().numpy(), re_ref) and np.array_equal(dim.cpu().numpy(), im_ref)


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("n", [0, 1, 7, 8, 9, 1000, 100003])
def test_convert_ragged_lengths(gpu_ctx, oracle, fmt, n):
    iq = oracle.synth_iq(fmt, max(n, 1))[: n * oracle.BYTES_PER_SAMPLE[fmt]]
    re, im = np.full(max(n, 1), 7.0, np.float32), np.full(max(n, 1), 7.0, np.float32)
    gpu_ctx.convert(fmt, iq if n else np.zeros(4, np.uint8), n, re, im)
    if n == 0:
        assert re[0] == 7.0  # untouched
        return
    re_ref, im_ref = oracle_fill(oracle, fmt, iq)
    assert np.array_equal(re[:n], re_ref) and np.array_equal(im[:n], im_ref)


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("fs,src,chan", [(2400000, 100_000_000, 100_250_000), (10_000_000, 433_000_000, 431_765_433),
                                         (20_000_000, 100_000_000, 100_000_000)])
def test_mix_bit_exact_across_packets(gpu_ctx, oracle, fmt, fs, src, chan):
    """The NCO index persists across packets (…java:119-126): three packets, each bit-exact."""
    L = oracle.lib()
    bps = oracle.BYTES_PER_SAMPLE[fmt]
    sizes = [8192, 1000, 4097]
    iq = oracle.synth_iq(fmt, sum(sizes))
    c = L.orc_converter_new(fmt)
    L.orc_converter_set_sample_rate(c, fs)
    L.orc_converter_set_frequency(c, src)
    eff, cos_t, sin_t = gpu_ctx.nco_design(fmt, fs, src - chan)
    idx, pos = 0, 0
    for n in sizes:
        chunk = iq[pos * bps:(pos + n) * bps].copy()
        sp = oracle.PacketView(n)
        assert L.orc_converter_mix(c, chunk, len(chunk), sp.p, chan) == n
        re, im = np.empty(n, np.float32), np.empty(n, np.float32)
        gpu_ctx.mix(fmt, chunk, n, cos_t, sin_t, idx, re, im)
        assert np.array_equal(re, sp.out_re()) and np.array_equal(im, sp.out_im())
        idx = (idx + n) % len(cos_t)
        assert idx == L.orc_converter_nco_index(c)
        pos += n
    assert eff == L.orc_converter_nco_freq(c)
    L.orc_converter_free(c)


def test_bad_arguments_are_reported(gpu_ctx):
    import rfanalyzer_b200 as rfa
    x = np.zeros(16, np.uint8)
    out = np.zeros(8, np.float32)
    with pytest.raises(rfa.RfaError):
        gpu_ctx.convert(9, x, 8, out, out)
    with pytest.raises(rfa.RfaError):
        gpu_ctx.mix(0, x, 8, np.zeros(0, np.float32), np.zeros(0, np.float32), 0, out, out)
