"""Parity at BASELINE.json's full sizes (the bench workloads themselves, device-resident through
vhp_visibility_batch_dev): closed-form results where the domain has them, the oracle on samples
of the batch elsewhere."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def env():
    import torch
    import visibility_heuristic_path_planner_b200 as vhp
    stream = torch.cuda.Stream(0)
    ctx = vhp.torch_context(0, stream)
    yield vhp, ctx, torch, stream
    ctx.close()


def run_dev(env, maps, src, smap, tdt):
    vhp, ctx, torch, stream = env
    dev = torch.device("cuda", 0)
    occ_t = torch.from_numpy(maps).to(dev)
    src_t = torch.from_numpy(src).to(dev)
    smap_t = None if smap is None else torch.from_numpy(smap).to(dev)
    out_t = torch.empty((len(src),) + maps.shape[1:], dtype=tdt, device=dev)
    with torch.cuda.stream(stream):
        ctx.visibility_batch_dev(occ_t, src_t, out_t, smap_t)
    stream.synchronize()
    ctx.synchronize()
    return out_t


def test_c2_full_batch_closed_form(env):
    """BASELINE configs[1]: 4096 sources on the empty 1000 x 1000 grid, fp32 store.  Every cell is
    lit (1.0) except the column x = 0 / row y = 0 the reference never writes when the source is
    not on them: checked for all 4096 fields on the device."""
    from bench import workload
    _, _, torch, _ = env
    maps, src, smap, _ = workload("c2", 0)
    out = run_dev(env, maps, src, smap, torch.float32)
    n, ny, nx = out.shape
    assert n == 4096 and (ny, nx) == (1000, 1000)
    sx = torch.from_numpy(src[:, 0].astype(np.int64)).to(out.device)
    sy = torch.from_numpy(src[:, 1].astype(np.int64)).to(out.device)
    assert bool(((out == 0) | (out == 1)).all())
    assert bool((out[:, 1:, 1:] == 1).all())                 # the interior is lit
    col0, row0 = out[:, :, 0], out[:, 0, :]                  # (n, ny), (n, nx)
    ys = torch.arange(ny, device=out.device)[None, :]
    xs = torch.arange(nx, device=out.device)[None, :]
    # x = 0: written only by a source with sx == 0 (and then not at y = 0 below the source)
    exp_col0 = (sx[:, None] == 0) & ((ys > 0) | (sy[:, None] == 0))
    exp_row0 = (sy[:, None] == 0) & ((xs > 0) | (sx[:, None] == 0))
    assert bool((col0 == exp_col0.float()).all())
    assert bool((row0 == exp_row0.float()).all())


def test_c2s_batch_sample_vs_oracle(env, oracle):
    """the obstacle workload of the bench (shipped settings.config environment, 4096 sources):
    16 fields spread over the batch equal the oracle bit for bit (fp64 store)"""
    from bench import workload
    _, _, torch, _ = env
    maps, src, smap, _ = workload("c2s", 0)
    pick = np.linspace(0, len(src) - 1, 16).astype(int)
    sub = np.ascontiguousarray(src[: pick[-1] + 1])
    # run a 512-pair prefix in fp64 (16 GB for all 4096) plus the picked pairs of the rest
    part = np.ascontiguousarray(np.concatenate([sub[:512], src[pick[pick >= 512]]]))
    out = run_dev(env, maps, part, None, torch.float64).cpu().numpy()
    for k, (sxk, syk) in enumerate(part):
        if k < 512 and k % 64:
            continue
        ref = oracle.compute_visibility(maps[0].astype(np.float64), int(sxk), int(syk))
        assert np.array_equal(out[k], ref), (k, sxk, syk)


def test_c4_batch_sample_vs_oracle(env, oracle):
    """BASELINE configs[3] shape: 1024 random 256 x 256 maps x 16 sources per GPU; every source
    of a 24-map sample equals the oracle bit for bit, and fp32 store == fp64 rounded once"""
    from bench import workload
    _, _, torch, _ = env
    maps, src, smap, _ = workload("c4", 0)
    out64 = run_dev(env, maps, src, smap, torch.float64)
    out32 = run_dev(env, maps, src, smap, torch.float32)
    assert bool((out64.float() == out32).all())
    for m in np.linspace(0, maps.shape[0] - 1, 24).astype(int):
        occ = maps[m].astype(np.float64)
        for k in np.flatnonzero(smap == m):
            ref = oracle.compute_visibility(occ, int(src[k, 0]), int(src[k, 1]))
            assert np.array_equal(out64[k].cpu().numpy(), ref), (m, k)
