"""CPU: the GEMM work schedule (gemm.cu TileSched / plan_gemm, read through the host-only b200_gemm_schedule) covers every
(tile, k-block) exactly once, cuts a streamed tile at most once, and orders the segments so that the ordered reduce-adds
of stream-K can never wait on work that is scheduled later (the determinism + no-deadlock argument of DESIGN.md §4.1)."""
import ctypes as C

import numpy as np
import pytest

from latte_b200 import _lib

EPI_BIAS, EPI_GATE_RESIDUAL = 0, 2
SHAPES = [(8192, 1152, 4608), (8192, 1152, 1152), (8192, 3456, 1152), (8192, 4608, 1152), (4096, 1152, 4608), (16384, 1152, 4608),
          (32768, 1152, 4608), (8192 + 128, 1152, 2048), (20480, 384, 4096), (200, 192, 192), (128, 128, 64), (1024 * 37, 320, 2560)]


def schedule(M, N, K, epi, bn=0, sms=148):
    lib = _lib.load()
    bn_o, pairs, sk = C.c_int(), C.c_int(), C.c_int()
    n = lib.b200_gemm_schedule(M, N, K, epi, bn, sms, C.byref(bn_o), C.byref(pairs), C.byref(sk), None, 0)
    assert n >= 0, _lib.last_error()
    seg = (C.c_int32 * (4 * n))()
    assert lib.b200_gemm_schedule(M, N, K, epi, bn, sms, None, None, None, seg, n) == n
    return bn_o.value, pairs.value, sk.value, np.frombuffer(seg, dtype=np.int32).reshape(n, 4).copy()


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("epi", [EPI_BIAS, EPI_GATE_RESIDUAL])
@pytest.mark.parametrize("bn", [0, 128, 192, 256])
@pytest.mark.parametrize("sms", [148, 132, 16])
def test_schedule_covers_and_orders(shape, epi, bn, sms):
    M, N, K = shape
    bn_used, pairs, sk, seg = schedule(M, N, K, epi, bn, sms)
    assert bn_used in (128, 192, 256) and (bn == 0 or bn_used == bn)
    num_kb = K // 64
    tiles = ((M + 127) // 128 + 1) // 2 * ((N + bn_used - 1) // bn_used)
    assert pairs == min(tiles, sms // 2)
    assert sk == 0 or epi == EPI_GATE_RESIDUAL          # partial sums only where they can be reduce-added
    # coverage: every k-block of every tile exactly once
    cover = np.zeros((tiles, num_kb), dtype=np.int32)
    for p, t, k0, k1 in seg:
        assert 0 <= p < pairs and 0 <= t < tiles and 0 <= k0 < k1 <= num_kb
        cover[t, k0:k1] += 1
    assert (cover == 1).all()
    # per tile: at most two segments (one cut), contiguous in k
    by_tile = {}
    for i, (p, t, k0, k1) in enumerate(seg):
        by_tile.setdefault(int(t), []).append((int(k0), int(k1), int(p), i))
    order_in_pair = {}
    for i, (p, t, k0, k1) in enumerate(seg):
        order_in_pair.setdefault(int(p), []).append(i)
    pos = {i: n for p, lst in order_in_pair.items() for n, i in enumerate(lst)}
    for t, parts in by_tile.items():
        parts.sort()
        assert len(parts) <= 2, f"tile {t} cut {len(parts) - 1} times"
        if len(parts) == 2:
            (a0, a1, pa, ia), (b0, b1, pb, ib) = parts
            assert sk == 1 and a0 == 0 and a1 == b0 and b1 == num_kb and pb == pa + 1
            # the continuing segment is the LAST thing its pair does; the starting segment is not after any k > 0 segment
            # of its own pair -> the wait in the continuing segment's epilogue is on work that never waits itself
            assert pos[ib] == len(order_in_pair[pb]) - 1
            assert all(seg[j][2] == 0 for j in order_in_pair[pa][:pos[ia] + 1])
    # balance: with stream-K no pair has more than one k-block above the mean share
    work = np.zeros(pairs, dtype=np.int64)
    for p, t, k0, k1 in seg:
        work[p] += k1 - k0
    if sk:
        assert work.max() - work.min() <= 1
    else:
        assert work.max() - work.min() <= num_kb


def test_latte_shapes_pick_the_measured_configuration():
    """XL/2 at B_model = 2 on 148 SMs: QKV, fc1 -> 256-wide tiles, data-parallel; proj -> 192, data-parallel
    (K too short to split); fc2 -> 256-wide, last two waves streamed along K."""
    assert schedule(8192, 3456, 1152, EPI_BIAS)[:3] == (256, 74, 0)
    assert schedule(8192, 4608, 1152, EPI_BIAS)[:3] == (256, 74, 0)
    assert schedule(8192, 1152, 1152, EPI_GATE_RESIDUAL)[:3] == (192, 74, 0)
    assert schedule(8192, 1152, 4608, EPI_GATE_RESIDUAL)[:3] == (256, 74, 1)


# ------------------------------------------------------------------------------------------------ weight gradients
def wgrad_schedule(rows, n_out, n_in, sms=148):
    lib = _lib.load()
    bn_o, pairs, sk = C.c_int(), C.c_int(), C.c_int()
    n = lib.b200_wgrad_schedule(rows, n_out, n_in, sms, C.byref(bn_o), C.byref(pairs), C.byref(sk), None, 0)
    assert n >= 0, _lib.last_error()
    seg = (C.c_int32 * (4 * n))()
    assert lib.b200_wgrad_schedule(rows, n_out, n_in, sms, None, None, None, seg, n) == n
    return bn_o.value, pairs.value, sk.value, np.frombuffer(seg, dtype=np.int32).reshape(n, 4).copy()


@pytest.mark.parametrize("shape", [(20480, 1152, 1152), (20480, 3456, 1152), (20480, 4608, 1152), (20480, 1152, 4608), (8192, 1152, 1152),
                                   (4096, 512, 256), (1536, 384, 128), (20480, 32, 1152), (2048, 1152, 1152)])
@pytest.mark.parametrize("sms", [148, 132, 16])
def test_wgrad_schedule_chains_cannot_deadlock(shape, sms):
    """b200_wgrad (dW[n_out, n_in] += dY^T X over `rows`): when the output has fewer tiles than CTA pairs every tile is cut into
    several K-segments.  The ordered reduce-adds make a segment with kb0 > 0 wait (in its epilogue) for the segment that ends at
    kb0; with every pair executing its segments in the listed order, a simulation must finish -- i.e. no cyclic wait -- and the
    segments of a tile must go to consecutive pairs in k order."""
    rows, n_out, n_in = shape
    bn, pairs, sk, seg = wgrad_schedule(rows, n_out, n_in, sms)
    num_kb = rows // 64
    tiles = ((n_out + 127) // 128 + 1) // 2 * ((n_in + bn - 1) // bn)
    assert bn in (128, 256) and 1 <= pairs <= sms // 2
    cover = np.zeros((tiles, num_kb), dtype=np.int32)
    for p, t, k0, k1 in seg:
        assert 0 <= p < pairs and 0 <= t < tiles and 0 <= k0 < k1 <= num_kb
        cover[t, k0:k1] += 1
    assert (cover == 1).all()
    by_tile = {}
    for p, t, k0, k1 in seg:
        by_tile.setdefault(int(t), []).append((int(k0), int(k1), int(p)))
    for t, parts in by_tile.items():
        parts.sort()
        if len(parts) > 1:
            assert sk == 1
            for (a0, a1, pa), (b0, b1, pb) in zip(parts, parts[1:]):
                assert a1 == b0 and pb == pa + 1            # contiguous in k, handed from one pair to the next
    if tiles < pairs:
        assert sk == 1 and pairs == sms // 2                # small outputs are spread over every pair
    # simulate: each pair runs its segments in order; a kb0 > 0 segment completes only after its predecessor in the tile did
    queues = {}
    for p, t, k0, k1 in seg:
        queues.setdefault(int(p), []).append((int(t), int(k0), int(k1)))
    done = set()                                            # (tile, kb1) of completed segments
    heads = {p: 0 for p in queues}
    progress = True
    while progress:
        progress = False
        for p, q in queues.items():
            while heads[p] < len(q):
                t, k0, k1 = q[heads[p]]
                # the MMAs of later segments may run ahead (two accumulators), but completion is in order: model the strictest case
                if k0 == 0 or (t, k0) in done:
                    done.add((t, k1))
                    heads[p] += 1
                    progress = True
                else:
                    break
    assert all(heads[p] == len(q) for p, q in queues.items()), "cyclic wait in the ordered stream-K chains"
