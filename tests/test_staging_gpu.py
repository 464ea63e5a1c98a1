"""Pageable host inputs of a few MB and more are staged by several host threads through pinned bounce buffers
(nds_stage.cu).  The result must not depend on how the bytes travelled: the same call on a device-resident copy, on
pinned memory and against the oracle on a sample of lanes."""
import numpy as np
import pytest


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2049, 4099), (4096, 4096), (3, 1 << 22)])
def test_pageable_input_matches_device_input(nds, gpu_ctx, oracle, shape):
    import torch
    rng = np.random.default_rng(shape[0])
    a = rng.random(shape, dtype=np.float32)          # 12 ... 64 MiB of ordinary (pageable) memory: not a multiple of a chunk
    assert a.nbytes >= 8 << 20
    qs = [0.25, 0.5, 0.75]
    got = nds.quantiles_axis_mut(a, 1, qs, nds.Linear, ctx=gpu_ctx)
    dev = nds.quantiles_axis_mut(torch.from_numpy(a).cuda(), 1, qs, nds.Linear, ctx=gpu_ctx)
    assert np.array_equal(got.view(np.uint32), dev.cpu().numpy().view(np.uint32))
    rows = np.unique(np.linspace(0, shape[0] - 1, 8).astype(int))
    st, want, _ = oracle.quantiles_axis(np.ascontiguousarray(a[rows]), 1, qs, "linear", select_impl=1)
    assert st == 0 and np.array_equal(got[rows].view(np.uint32), want.view(np.uint32))
    pinned = torch.from_numpy(a).pin_memory().numpy()
    again = nds.quantiles_axis_mut(pinned, 1, qs, nds.Linear, ctx=gpu_ctx)
    assert np.array_equal(got.view(np.uint32), again.view(np.uint32))


@pytest.mark.gpu
def test_pageable_strided_view_and_repeated_calls(nds, gpu_ctx, oracle):
    rng = np.random.default_rng(3)
    base = rng.standard_normal((3000, 2500))          # 60 MB f64
    view = base[::-1, 3:2400:2]                       # negative row stride, strided columns: the span is staged as it lies
    for _ in range(3):                                # the bounce buffers are reused from call to call
        got = nds.quantiles_axis_mut(view, 0, [0.1, 0.9], nds.Midpoint, ctx=gpu_ctx)
        st, want, _ = oracle.quantiles_axis(np.ascontiguousarray(view[:, :16]), 0, [0.1, 0.9], "midpoint", select_impl=1)
        assert st == 0 and np.array_equal(got[:, :16].view(np.uint64), want.view(np.uint64))
        base[0, 0] += 1.0                             # the next call must see the new bytes, not a stale buffer
