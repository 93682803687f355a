"""Host half of the packed result transport (csrc/host_expand.cpp) without a GPU: a numpy
restatement of the device-side packing (result_transport.cu) produces the chunk, the
library's thread pool expands it, the bytes must come back exactly."""
import ctypes as C

import numpy as np
import pytest

UNIT = 128


@pytest.fixture(scope="module")
def lib():
    import visibility_heuristic_path_planner_b200 as vhp
    return vhp.load_library()


def pack_numpy(raw: bytes, elem: int):
    """uniform (all 0.0 or all 1.0) / literal classification of 128-byte units, as
    pack_results_kernel does it; returns (mask, word_base, vmask, literals, nunits, nlit)"""
    valid = len(raw)
    nunits = (valid + UNIT - 1) // UNIT
    nwords = (nunits + 31) // 32
    buf = np.zeros(nunits * UNIT, np.uint8)
    buf[:valid] = np.frombuffer(raw, np.uint8)
    buf[valid:] = 0xAB  # whatever lies behind the chunk on the device
    units = buf.reshape(nunits, UNIT)
    mask = np.zeros(nwords, np.uint32)
    base = np.zeros(nwords, np.uint32)
    vmask = np.zeros(nwords, np.uint32)
    one = np.array([1.0], np.float32 if elem == 4 else np.float64).view(np.uint32 if elem == 4 else np.uint64)[0]
    lits, cursor = [], 0
    # words in a scrambled order, like warps racing for the cursor
    order = np.random.default_rng(7).permutation(nwords)
    slots = {}
    for w in order:
        m, vm, mine = 0, 0, []
        for u in range(32):
            k = w * 32 + u
            if k >= nunits:
                break
            e = units[k].view(np.uint32 if elem == 4 else np.uint64)
            if (e == e[0]).all() and e[0] in (0, one):
                vm |= (1 << u) if e[0] == one else 0
            else:
                m |= 1 << u
                mine.append(units[k])
        mask[w] = m
        vmask[w] = vm
        base[w] = cursor
        slots[int(w)] = (cursor, mine)
        cursor += len(mine)
    lit = np.zeros(max(cursor, 1) * UNIT, np.uint8)
    for w, (b0, mine) in slots.items():
        for i, u in enumerate(mine):
            lit[(b0 + i) * UNIT:(b0 + i + 1) * UNIT] = u
    return mask, base, vmask, lit, nunits, cursor


@pytest.mark.parametrize("elem,dtype", [(4, np.float32), (8, np.float64)])
@pytest.mark.parametrize("n,offset", [(1, 0), (127, 0), (128, 0), (5000, 0), (70001, 0), (70001, 4), (33000, 8)])
@pytest.mark.parametrize("threads", [1, 5])
def test_expand_restores_the_bytes(lib, elem, dtype, n, offset, threads):
    g = np.random.default_rng(n + elem)
    a = np.ones(n, dtype)
    # flat runs of 1 and 0 with a few ramps in between, like a visibility field
    for _ in range(max(1, n // 700)):
        i, l = int(g.integers(0, n)), int(g.integers(1, 900))
        kind = int(g.integers(0, 3))
        a[i:i + l] = 0 if kind == 0 else (1 if kind == 1 else g.random(len(a[i:i + l])))
    raw = a.tobytes()
    mask, base, vmask, lit, nunits, nlit = pack_numpy(raw, elem)
    # destination with a chosen misalignment (offset 0: the non-temporal path)
    backing = np.full(len(raw) + 64 + 16, 0xEE, np.uint8)
    start = (-backing.ctypes.data) % 16 + offset
    dst = backing[start:start + len(raw)]
    st = lib.vhp_expand_packed_chunk(mask.ctypes.data, base.ctypes.data, vmask.ctypes.data, elem,
                                     lit.ctypes.data, nunits, len(raw), dst.ctypes.data, threads)
    assert st == 0
    assert dst.tobytes() == raw
    assert (backing[:start] == 0xEE).all() and (backing[start + len(raw):] == 0xEE).all()
    if n >= 5000:
        assert nlit < nunits  # something was uniform


def test_expand_rejects_inconsistent_sizes(lib):
    z = np.zeros(64, np.uint64)
    assert lib.vhp_expand_packed_chunk(z.ctypes.data, z.ctypes.data, z.ctypes.data, 8, z.ctypes.data,
                                       2, 100, z.ctypes.data, 1) != 0
    assert lib.vhp_expand_packed_chunk(z.ctypes.data, z.ctypes.data, z.ctypes.data, 3, z.ctypes.data,
                                       1, 100, z.ctypes.data, 1) != 0
    assert lib.vhp_expand_packed_chunk(None, None, None, 4, None, 0, 0, None, 1) == 0


@pytest.mark.parametrize("elem,dtype", [(4, np.float32), (8, np.float64)])
def test_direct_mode_leaves_literal_units_alone(lib, elem, dtype):
    """literals == NULL: the device has stored the literal units in dst; the host threads write
    the uniform units only"""
    g = np.random.default_rng(5)
    n = 40000
    a = np.ones(n, dtype)
    for _ in range(60):
        i, l = int(g.integers(0, n)), int(g.integers(1, 700))
        a[i:i + l] = g.random(len(a[i:i + l])) if g.random() < 0.5 else 0
    raw = a.tobytes()
    assert len(raw) % UNIT == 0
    mask, base, vmask, lit, nunits, nlit = pack_numpy(raw, elem)
    dst = np.full(len(raw), 0xEE, np.uint8)
    litmask = np.zeros(nunits, bool)
    for w in range(len(mask)):
        for u in range(32):
            if (int(mask[w]) >> u) & 1:
                litmask[w * 32 + u] = True
    src = np.frombuffer(raw, np.uint8).reshape(nunits, UNIT)
    view = dst.reshape(nunits, UNIT)
    view[litmask] = src[litmask]          # what the device's stores leave behind
    st = lib.vhp_expand_packed_chunk(mask.ctypes.data, base.ctypes.data, vmask.ctypes.data, elem,
                                     None, nunits, len(raw), dst.ctypes.data, 3)
    assert st == 0
    assert dst.tobytes() == raw
    assert 0 < nlit < nunits


def test_expand_fuzz(lib):
    """random byte strings of every length class: all-literal, all-uniform, partial last units"""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(1, 9000), st.sampled_from([4, 8]), st.integers(0, 3), st.integers(0, 2 ** 31))
    def run(nelem, elem, kind, seed):
        g = np.random.default_rng(seed)
        dt = np.uint32 if elem == 4 else np.uint64
        if kind == 0:
            a = g.integers(0, 2 ** 31, nelem).astype(dt)                       # nothing compresses
        elif kind == 1:
            a = np.full(nelem, 0x3F800000 if elem == 4 else 0x3FF0000000000000, dt)  # all lit
        else:
            a = np.repeat(g.integers(0, 3, nelem // 40 + 1), 40)[:nelem].astype(dt)  # flat runs
        raw = a.tobytes()
        mask, base, vmask, lit, nunits, _ = pack_numpy(raw, elem)
        dst = np.full(len(raw) + 32, 0xEE, np.uint8)
        off = (-dst.ctypes.data) % 16
        d = dst[off:off + len(raw)]
        assert lib.vhp_expand_packed_chunk(mask.ctypes.data, base.ctypes.data, vmask.ctypes.data, elem,
                                           lit.ctypes.data, nunits, len(raw), d.ctypes.data, 2) == 0
        assert d.tobytes() == raw
        assert (dst[off + len(raw):] == 0xEE).all()

    run()


def test_expand_range_fuzz(lib):
    """vhp_expand_packed_range (the lazy expansion behind vhp_packed_expand): any byte range that starts
    on an element boundary -- inside units, across mask words, up to a partial last unit -- equals that
    slice of the original bytes, and nothing outside the destination is touched."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=80, deadline=None)
    @given(st.integers(1, 12000), st.sampled_from([4, 8]), st.integers(0, 2), st.integers(0, 2 ** 31))
    def run(nelem, elem, kind, seed):
        g = np.random.default_rng(seed)
        dt = np.float32 if elem == 4 else np.float64
        a = np.ones(nelem, dt)
        if kind == 0:
            a = g.random(nelem).astype(dt)                                        # nothing compresses
        elif kind == 1:
            for _ in range(max(1, nelem // 300)):                                # flat runs and ramps
                i, l = int(g.integers(0, nelem)), int(g.integers(1, 500))
                k = int(g.integers(0, 4))
                a[i:i + l] = (0, 1, 0.25)[k] if k < 3 else g.random(len(a[i:i + l]))
        raw = a.tobytes()
        mask, base, vmask, lit, nunits, _ = pack_numpy(raw, elem)
        for _ in range(6):
            e0 = int(g.integers(0, nelem))
            e1 = int(g.integers(e0, nelem + 1))
            b0, b1 = e0 * elem, e1 * elem
            dst = np.full(b1 - b0 + 48, 0xEE, np.uint8)
            off = 16 + int(g.integers(0, 2)) * 4
            assert lib.vhp_expand_packed_range(mask.ctypes.data, base.ctypes.data, vmask.ctypes.data, elem,
                                               lit.ctypes.data, nunits, len(raw), b0, b1,
                                               dst[off:].ctypes.data) == 0
            assert dst[off:off + b1 - b0].tobytes() == raw[b0:b1]
            assert (dst[:off] == 0xEE).all() and (dst[off + b1 - b0:] == 0xEE).all()

    run()
    z = np.zeros(64, np.uint32)
    assert lib.vhp_expand_packed_range(z.ctypes.data, z.ctypes.data, z.ctypes.data, 4, z.ctypes.data, 1, 100, 2, 50,
                                       z.ctypes.data) != 0   # not on an element boundary
    assert lib.vhp_expand_packed_range(z.ctypes.data, z.ctypes.data, z.ctypes.data, 4, z.ctypes.data, 1, 100, 0, 104,
                                       z.ctypes.data) != 0   # past the end
