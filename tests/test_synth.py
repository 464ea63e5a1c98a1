"""The synthetic input streams of the benchmark: the numpy restatement is self-consistent on CPU, and
the device generator (`nds_synth_fill`) writes the same bytes (GPU)."""
import numpy as np
import pytest


def test_streams_are_counter_based_and_shaped(nds):
    s = nds.synth
    # element i depends on (seed + i) only: any sub-range equals the slice of a longer range
    for kind in s.KINDS:
        a = s.host(kind, 1000, 1234)
        b = s.host(kind, 300, 1234, first_index=500)
        assert a.dtype == np.dtype(s.DTYPES[s.KINDS[kind]])
        assert np.array_equal(a[500:800].view(np.uint8), b.view(np.uint8))
    # splitmix64 known answers (state 0: first outputs of the reference generator)
    z = s.splitmix64(np.array([0, 0x9E3779B97F4A7C15], dtype=np.uint64))
    assert int(z[0]) == 0xE220A8397B1DCDAF and int(z[1]) == 0x6E789E6AA1B965F4
    u = s.host("f64_uniform", 100000, 42)
    assert 0.0 <= u.min() and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.01
    f = s.host("f32_uniform_nan10", 200000, 1003)
    assert abs(np.isnan(f).mean() - 0.1) < 0.01
    assert np.all(f[~np.isnan(f)] < 1.0)
    assert set(np.unique(s.host("u32_16values", 5000, 1005)).tolist()) == {k * 0x11111111 for k in range(16)}
    assert set(np.unique(s.host("u32_16values_adv", 5000, 1005)).tolist()) == {0x80000000 + k for k in range(16)}
    assert set(np.unique(s.host("i64_16values", 5000, 1005)).tolist()) == {(k - 8) << 40 for k in range(16)}
    # host_at: column lanes of a C-order matrix
    idx = np.arange(4, dtype=np.uint64)[:, None] * np.uint64(10) + np.arange(3, dtype=np.uint64)[None, :]
    full = s.host("f32_uniform", 40, 7).reshape(4, 10)
    assert np.array_equal(s.host_at("f32_uniform", idx, 7), full[:, :3])


@pytest.mark.gpu
def test_device_fill_matches_host(nds, gpu_ctx):
    import torch
    s = nds.synth
    tdt = {np.float64: torch.float64, np.float32: torch.float32, np.uint32: torch.uint32, np.int64: torch.int64}
    for kind, code in s.KINDS.items():
        n = 100003
        t = torch.empty(n, dtype=tdt[s.DTYPES[code]], device="cuda")
        s.fill(gpu_ctx, t, kind, 987654321, first_index=17)
        gpu_ctx.synchronize()
        got = t.cpu().numpy()
        want = s.host(kind, n, 987654321, first_index=17)
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8)), kind
