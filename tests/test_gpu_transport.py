"""Result transport of the host-buffer entry points (vhp_visibility_batch, vhp_raycast_batch):
the packed transport (uniform / literal 128-byte units, expanded by host threads) must leave
exactly the bytes of the plain device-to-host copy in the caller's buffer, and those must
equal the oracle."""
import numpy as np
import pytest

from conftest import rect_map

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def vhp():
    import visibility_heuristic_path_planner_b200 as m
    return m


@pytest.fixture(scope="module")
def ctx(vhp):
    c = vhp.Context(0)
    yield c
    c.close()


def sources(occ, n, seed):
    g = np.random.default_rng(seed)
    free = np.argwhere(occ != 0)
    pick = free[g.integers(0, len(free), n)]
    return np.ascontiguousarray(pick[:, ::-1]).astype(np.int32)


@pytest.mark.parametrize("dtype_name", ["F32", "F64"])
@pytest.mark.parametrize("shape", [(300, 217), (101, 101), (64, 64), (1, 9), (37, 1)])
def test_packed_equals_plain_and_oracle(ctx, vhp, oracle, shape, dtype_name):
    nx, ny = shape
    dtype = getattr(vhp, dtype_name)
    occ = rect_map(nx, ny, max(1, nx * ny // 1500), seed=nx + ny, lo=3, hi=25)
    occ[0, 0] = 1
    srcs = sources(occ, 9, seed=3)
    ctx.set_result_transport(0)
    plain = ctx.visibility_batch(occ, srcs, dtype=dtype)
    assert ctx.last_transport()[2] == 0
    ctx.set_result_transport(2)
    packed = ctx.visibility_batch(occ, srcs, dtype=dtype)
    d2h, res, was_packed = ctx.last_transport()
    assert was_packed == 1 and res == plain.nbytes  # pageable buffer: staged literal stream
    assert packed.dtype == plain.dtype and packed.tobytes() == plain.tobytes()
    for s, v in zip(srcs[:3], packed):
        ref = oracle.compute_visibility(occ, int(s[0]), int(s[1]))
        assert np.array_equal(v, ref if dtype == vhp.F64 else ref.astype(np.float32))
    ctx.set_result_transport(1)


def test_packed_raycast_equals_plain(ctx, vhp):
    occ = rect_map(200, 150, 20, seed=5, lo=3, hi=15)
    srcs = sources(occ, 4, seed=1)
    ctx.set_result_transport(0)
    plain = ctx.raycast_batch(occ, srcs, dtype=vhp.F32)
    ctx.set_result_transport(2)
    packed = ctx.raycast_batch(occ, srcs, dtype=vhp.F32)
    assert packed.tobytes() == plain.tobytes()
    ctx.set_result_transport(1)


def test_many_chunks_and_compression_1000(ctx, vhp):
    """several 256 MB chunks through the three buffer sets; flat fields shrink on the wire"""
    nx = ny = 1000
    g = np.random.default_rng(11)
    occ = np.ones((ny, nx), np.uint8)
    for _ in range(15):
        x, y = int(g.integers(1, nx)), int(g.integers(1, ny))
        w, h = int(g.integers(100, 201)), int(g.integers(100, 201))
        occ[y:y + h, x:x + w] = 0
    srcs = sources(occ, 300, seed=2)  # 1.2 GB of fp32 results: 5 chunks
    ctx.set_result_transport(0)
    plain = ctx.visibility_batch(occ, srcs, dtype=vhp.F32)
    d2h_plain, res, _ = ctx.last_transport()
    assert d2h_plain == res == plain.nbytes
    ctx.set_result_transport(1)  # automatic: large and compressible -> packed
    packed = ctx.visibility_batch(occ, srcs, dtype=vhp.F32)
    d2h, res2, was_packed = ctx.last_transport()
    assert was_packed and res2 == res
    assert d2h < 0.7 * res, (d2h, res)
    assert np.array_equal(packed, plain)


def test_automatic_mode_falls_back_when_results_do_not_compress(ctx, vhp):
    """isolated single-cell obstacles put every unit into the penumbra: after the first chunks
    the call continues with plain copies; the result is the same"""
    nx = ny = 1000
    g = np.random.default_rng(13)
    occ = (g.random((ny, nx)) > 0.004).astype(np.uint8)
    srcs = sources(occ, 280, seed=4)
    ctx.set_result_transport(0)
    plain = ctx.visibility_batch(occ, srcs, dtype=vhp.F32)
    ctx.set_result_transport(1)
    auto = ctx.visibility_batch(occ, srcs, dtype=vhp.F32)
    d2h, res, was_packed = ctx.last_transport()
    assert np.array_equal(auto, plain)
    assert was_packed and d2h > 0.6 * res  # tried, measured, gave up
    ctx.set_result_transport(2)
    forced = ctx.visibility_batch(occ, srcs, dtype=vhp.F32)
    assert np.array_equal(forced, plain)
    ctx.set_result_transport(1)


@pytest.mark.parametrize("gpu_share", [0, 5, 16])
@pytest.mark.parametrize("dtype_name", ["F32", "F64"])
@pytest.mark.parametrize("shape,npairs", [((100, 101), 3), ((300, 217), 9), ((1000, 1000), 150)])
def test_direct_mode_into_pinned_memory(ctx, vhp, shape, npairs, dtype_name, gpu_share):
    """a pinned caller buffer: the device stores the literal units straight into it, the host
    threads write the uniform ones (a partial last unit comes through the meta block)"""
    import torch
    nx, ny = shape
    dtype = getattr(vhp, dtype_name)
    if nx == 1000 and dtype == vhp.F64:
        pytest.skip("one large case is enough")
    occ = rect_map(nx, ny, max(2, nx * ny // 3000), seed=nx * 3 + ny, lo=3, hi=40)
    srcs = sources(occ, npairs, seed=8)
    ctx.set_result_transport(0)
    plain = ctx.visibility_batch(occ, srcs, dtype=dtype)
    tdt = torch.float32 if dtype == vhp.F32 else torch.float64
    pinned = torch.full((npairs, ny, nx), float("nan"), dtype=tdt).pin_memory()
    ctx.set_result_transport(2)
    ctx.set_result_gpu_share(gpu_share)
    got = ctx.visibility_batch(occ, srcs, dtype=dtype, out=pinned.numpy())
    d2h, res, mode = ctx.last_transport()
    assert mode == 2 and res == plain.nbytes
    if gpu_share == 16:
        assert d2h >= 0.9 * res  # (nearly) everything came from the device
    assert got.tobytes() == plain.tobytes()
    # a second call must rewrite every byte as well
    pinned.fill_(float("nan"))
    ctx.visibility_batch(occ, srcs, dtype=dtype, out=pinned.numpy())
    assert pinned.numpy().tobytes() == plain.tobytes()
    ctx.set_result_transport(1)
    ctx.set_result_gpu_share(0)
