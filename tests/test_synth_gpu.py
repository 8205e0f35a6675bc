"""The device-side synthetic IQ generator is bit-identical to the oracle's (so full-size inputs
generated in place on the GPU can be spot-checked against the CPU path)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("first,n", [(0, 4096), (123457, 10001), (2 ** 33 + 5, 777), (7, 1)])
def test_device_generator_matches_oracle(gpu_ctx, oracle, fmt, first, n):
    import rfanalyzer_b200 as rfa
    want = oracle.synth_iq(fmt, n, first=first)
    got = np.zeros_like(want)
    rfa.synth_iq(gpu_ctx, fmt, n, got, first=first)
    assert np.array_equal(got, want)


def test_fm_component_matches_oracle(gpu_ctx, oracle):
    import rfanalyzer_b200 as rfa
    comps = [(rfa.synth_step(250000 / 2400000), 60, rfa.synth_step(1000 / 2400000), 3130000)]
    want = oracle.synth_iq(1, 50000, comps=comps, noise_shift=4)
    got = np.zeros_like(want)
    rfa.synth_iq(gpu_ctx, 1, 50000, got, comps=comps, noise_shift=4)
    assert np.array_equal(got, want)
    assert rfa.default_synth_components(0) == oracle.default_comps(0)
    assert rfa.default_synth_components(2) == oracle.default_comps(2)
