"""Argument guards of the CUDA engine (round-1 review): game ids outside the handle, select after an expand that dropped rows,
host buffers that are pinned but not naturally aligned.  Nothing here may fault the device: every case ends with a healthy
synchronize and an unchanged neighbour game."""
import ctypes as C

import numpy as np
import pytest

import harness

pytestmark = pytest.mark.gpu


def _fresh(n=200, **kw):
    from b200tetris import Engine
    e = Engine(n, 22, 10, piece_stream_len=256, hole_stream_len=64, **kw)
    e.reset(harness.seeds_for(np.arange(n), np.zeros(n, int)))
    return e


def test_host_calls_refuse_ids_outside_the_handle():
    from b200tetris import B2TError
    e = _fresh()
    before = e.export_states()
    for bad in ([0, 200], [-1], [5, 1 << 30]):
        idx = np.array(bad, np.int32)
        n = len(bad)
        # through the Engine wrapper: IndexError before the library is called (like the reference's list indexing)
        with pytest.raises(IndexError):
            e.reset(np.zeros(n, np.uint32), bad)
        with pytest.raises(IndexError):
            e.step(np.zeros(n, np.uint8), np.zeros((n, 16), np.uint8), np.ones(n, np.uint8), bad)
        with pytest.raises(IndexError):
            e.observe(bad)
        with pytest.raises(IndexError):
            e.enumerate(np.zeros(n, np.uint8), bad)
        with pytest.raises(IndexError):
            e.expand(np.zeros(n, np.uint8), bad)
        with pytest.raises(IndexError):
            e.export_states(bad)
        # straight at the C ABI: B2T_ERR_ARG (-1), no launch
        ip = idx.ctypes.data_as(C.POINTER(C.c_int32))
        z8 = np.zeros(n * 16, np.uint8); zp = z8.ctypes.data_as(C.POINTER(C.c_uint8))
        z32 = np.zeros(n, np.uint32)
        assert e.lib.b2t_reset_host(e._h, ip, n, z32.ctypes.data_as(C.POINTER(C.c_uint32))) == -1
        assert e.lib.b2t_step_host(e._h, ip, n, zp, zp, zp, 16, None, 1, 0, None, None, None, None, None, None, None, None, 0, 0) == -1
        assert e.lib.b2t_observe_host(e._h, ip, n, None, None, None, None) == -1
        assert e.lib.b2t_enumerate_host(e._h, ip, n, zp, None, None, None, 0, 0) == -1
        buf = np.zeros((n, e.state_bytes), np.uint8)
        assert e.lib.b2t_export_states(e._h, ip, n, buf.ctypes.data_as(C.c_void_p)) == -1
        assert e.lib.b2t_import_states(e._h, ip, n, buf.ctypes.data_as(C.c_void_p)) == -1
        assert b"game id" in e.lib.b2t_last_error(e._h)
    assert np.array_equal(before, e.export_states())
    e.close()


def test_device_calls_treat_ids_outside_the_handle_as_inert_rows():
    import torch
    e = _fresh()
    dev = torch.device("cuda:0")
    n = e.n
    before = e.export_states()
    idx = torch.tensor([3, -1, 200, 1 << 30, -(1 << 31), 7], dtype=torch.int32, device=dev)
    m = idx.numel()
    pl = torch.zeros(m, dtype=torch.uint8, device=dev)
    keys = torch.zeros((m, 16), dtype=torch.uint8, device=dev); keys[:, 0] = 3; keys[:, 1] = 7
    klen = torch.full((m,), 2, dtype=torch.uint8, device=dev)
    field = torch.full((m, 2, 22, 10), 9, dtype=torch.uint8, device=dev); vec = torch.full((m, 2, 19), 9, dtype=torch.uint8, device=dev)
    done = torch.full((m,), 9, dtype=torch.uint8, device=dev)
    e.step_dev(pl, keys, klen, idx=idx, done=done, field=field, vec=vec)
    torch.cuda.synchronize()
    after = e.export_states()
    changed = [g for g in range(n) if not np.array_equal(before[g], after[g])]
    assert changed == [3, 7]
    assert field[1:5].sum().item() == 0 and vec[1:5].sum().item() == 0          # inert rows: zero observations
    assert done.cpu().tolist()[1:5] == [9, 9, 9, 9]                               # ... and untouched results
    assert field[0].sum().item() > 0 and field[5].sum().item() > 0
    # enumerate / expand: K = 0 for the inert rows, nothing written for them
    K = torch.full((m,), -5, dtype=torch.int32, device=dev); offs = torch.empty(m + 1, dtype=torch.int32, device=dev)
    cap = m * e.kmax
    pf = torch.zeros((cap, 2, 22, 10), dtype=torch.uint8, device=dev)
    e.expand_dev(pl, K, offs, cap, idx=idx, field=pf)
    torch.cuda.synchronize()
    k = K.cpu().tolist()
    assert k[1:5] == [0, 0, 0, 0] and k[0] > 0 and k[5] > 0 and offs[-1].item() == k[0] + k[5]
    # reset / observe with a bad id
    e.reset_dev(torch.zeros(m, dtype=torch.int32, device=dev), idx)
    e.observe_dev(idx=idx, field=field)
    torch.cuda.synchronize()
    assert np.array_equal(e.export_states([0, 1, 2, 4, 5, 6, 199]), after[[0, 1, 2, 4, 5, 6, 199]])
    e.close()


def test_select_never_reads_rows_the_expand_dropped():
    import torch
    e = _fresh(64)
    dev = torch.device("cuda:0")
    n = e.n
    pl = torch.zeros(n, dtype=torch.uint8, device=dev)
    K = torch.empty(n, dtype=torch.int32, device=dev); offs = torch.empty(n + 1, dtype=torch.int32, device=dev)
    full = n * e.kmax
    e.expand_dev(pl, K, offs, full)
    torch.cuda.synchronize()
    total = int(offs[-1].item())
    cap = total // 2 + 3                                    # an expand that keeps about half of the rows
    e.expand_dev(pl, K, offs, cap)
    # scores: the caller's buffer has exactly `cap` rows; a guard page of +inf right behind it would win every argmax if it were read
    scores = torch.full((full,), float("inf"), device=dev)
    scores[:cap] = torch.rand(cap, device=dev)
    sel = torch.full((n,), -7, dtype=torch.int32, device=dev)
    e.select_dev(scores, offs, sel, cap)
    torch.cuda.synchronize()
    o = offs.cpu().numpy(); s = scores.cpu().numpy(); got = sel.cpu().numpy()
    for i in range(n):
        a, b = int(o[i]), min(int(o[i + 1]), cap)
        want = int(np.argmax(s[a:b])) if b > a else 0
        assert got[i] == want, (i, a, b)
    assert (got >= 0).all() and all(o[i] + got[i] < cap or got[i] == 0 for i in range(n))
    e.close()


def test_pinned_but_misaligned_host_buffers_take_the_staged_path():
    """b2t_step_host with every buffer pinned normally runs zero-copy (the kernel dereferences the caller's pages); a sliced pinned
    array is not naturally aligned for the kernel's 32-bit / 16-bit accesses and must go through staging with identical results."""
    from b200tetris import Engine, host_array
    from b200tetris.engine import _hp
    n = 1 << 16                                              # PIPE_MIN_ROWS: the zero-copy / pipelined regime
    gids = np.arange(n)
    seeds = harness.seeds_for(gids, np.zeros(n, int))
    outs = []
    for shift in (0, 1):                                     # 0: aligned (zero-copy), 1: every array starts one byte into its allocation
        e = Engine(n, 22, 10, piece_stream_len=256, hole_stream_len=64)
        e.reset(seeds)
        def arr(count, dtype):
            raw = host_array(count * np.dtype(dtype).itemsize + 8, np.uint8)
            return raw[shift:shift + count * np.dtype(dtype).itemsize].view(dtype) if shift == 0 else \
                np.frombuffer(raw, np.uint8, count * np.dtype(dtype).itemsize, shift)
        player = arr(n, np.uint8); keys = arr(n * 16, np.uint8); klen = arr(n, np.uint8)
        rs_b = arr(n, np.uint32); rew = arr(n, np.int8); done = arr(n, np.uint8); lines = arr(2 * n, np.uint8)
        from bench import sventon_keys
        res = []
        for t in range(6):
            pl = ((gids + t) & 1).astype(np.uint8)
            k, ln = sventon_keys(gids, t)
            player[:] = pl; keys[:] = k.ravel(); klen[:] = ln
            rs = harness.seeds_for(gids, np.full(n, t + 1)).astype(np.uint32)
            rs_b.view(np.uint8)[:] = rs.view(np.uint8)
            rc = e.lib.b2t_step_host(e._h, None, n, _hp(player, C.c_uint8), _hp(keys, C.c_uint8), _hp(klen, C.c_uint8), 16, None, 1, 0,
                                     C.cast(rs_b.ctypes.data, C.POINTER(C.c_uint32)), C.cast(rew.ctypes.data, C.POINTER(C.c_int8)),
                                     _hp(done, C.c_uint8), _hp(lines, C.c_uint8), None, None, None, None, 0, 0)
            assert rc == 0, e.lib.b2t_last_error(e._h)
            res.append((np.array(rew.view(np.int8)), np.array(done), np.array(lines)))
        outs.append((res, e.export_states(np.arange(0, n, 257))))
        e.close()
    for (ra, rb) in zip(outs[0][0], outs[1][0]):
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
    assert np.array_equal(outs[0][1], outs[1][1])


def test_traffic_probe_runs_and_only_scribbles_on_its_own_handle():
    """b2t_debug_traffic (measurement aid of bench.py): B2T_OK, argument errors reported, the device stays healthy."""
    import torch
    e = _fresh(4096)
    dev = torch.device("cuda:0")
    n = e.n
    P = lambda t: C.c_void_p(t.data_ptr())
    keys = torch.zeros((n, 16), dtype=torch.uint8, device=dev); b1 = torch.zeros(n, dtype=torch.uint8, device=dev)
    seeds = torch.zeros(n, dtype=torch.int32, device=dev); rew = torch.empty(n, dtype=torch.int8, device=dev)
    lines = torch.empty((n, 2), dtype=torch.uint8, device=dev); pidx = torch.empty((n, 2), dtype=torch.uint8, device=dev)
    field = torch.empty((n, 2, 22, 10), dtype=torch.uint8, device=dev); vec = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
    for mode in (1, 0):
        assert e.lib.b2t_debug_traffic(e._h, n, P(keys), 16, P(b1), P(b1), P(seeds), P(rew), P(b1.clone()), P(lines), P(field), P(vec), P(pidx), mode, None) == 0
    assert e.lib.b2t_debug_traffic(e._h, n + 1, P(keys), 16, P(b1), P(b1), P(seeds), P(rew), P(b1), P(lines), P(field), P(vec), P(pidx), 1, None) == -1
    assert e.lib.b2t_debug_traffic(e._h, n, None, 16, P(b1), P(b1), P(seeds), P(rew), P(b1), P(lines), P(field), P(vec), P(pidx), 1, None) == -1
    torch.cuda.synchronize()
    e.reset(harness.seeds_for(np.arange(n), np.zeros(n, int)))      # the handle is usable again after a reset
    K, _, _ = e.enumerate(np.zeros(n, np.uint8), want_keys=False)
    assert (K > 1).all()
    e.close()


@pytest.mark.parametrize("H,W,pieces", [(22, 10, (0, 1, 2, 3, 4, 5, 6)), (16, 8, (0, 6)), (40, 20, (0, 1, 2, 3, 4, 5, 6))])
def test_bit_packed_fields_equal_the_uint8_fields(H, W, pieces):
    """b2t_obs.dtype = B2T_BITS (an extra format): bit r*W+c of a player's SW words == the reference-shaped uint8 cell, from observe,
    from the fused step (hot variant), from a subset step (generic variant) and from the world model, in both player orders."""
    import torch
    from b200tetris import Engine
    dev = torch.device("cuda:0")
    n = 777
    e = Engine(n, H, W, pieces, piece_stream_len=256, hole_stream_len=64)
    gids = np.arange(n)
    e.reset(harness.seeds_for(gids, np.zeros(n, int)))
    SW = (H * W + 31) // 32
    pol = harness.Policy("junk", lmax=16)

    def unpack(bits):                     # [rows, 2, SW] int32 -> [rows, 2, H, W] uint8
        w = bits.cpu().numpy().view(np.uint32)
        b = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).astype(np.uint8).reshape(w.shape[0], 2, SW * 32)
        assert (b[..., H * W:] == 0).all()
        return b[..., :H * W].reshape(-1, 2, H, W)
    K = torch.empty(n, dtype=torch.int32, device=dev); offs = torch.empty(n + 1, dtype=torch.int32, device=dev)
    for t in range(25):
        player = ((gids + t) & 1).astype(np.uint8)
        keys, klen, _ = pol.choose(None, t, gids, player)
        pl = torch.from_numpy(player).to(dev); kd = torch.from_numpy(keys).to(dev); ld = torch.from_numpy(klen).to(dev)
        seeds = torch.from_numpy(harness.seeds_for(gids, np.full(n, t + 1)).astype(np.int64)).to(dev).to(torch.int32)
        fb = torch.full((n, 2, SW), -1, dtype=torch.int32, device=dev); vb = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
        persp = bool(t & 1)
        e.step_dev(pl, kd, ld, reset_seeds=seeds, field=fb, vec=vb, perspective=persp)          # hot variant, bits
        f8 = torch.empty((n, 2, H, W), dtype=torch.uint8, device=dev); v8 = torch.empty((n, 2, 19), dtype=torch.uint8, device=dev)
        e.observe_dev(field=f8, vec=v8, me=pl if persp else None)
        fb2 = torch.full((n, 2, SW), -1, dtype=torch.int32, device=dev)
        e.observe_dev(field=fb2, me=pl if persp else None)                                       # generic variant, bits
        torch.cuda.synchronize()
        assert np.array_equal(unpack(fb), f8.cpu().numpy()) and np.array_equal(unpack(fb2), f8.cpu().numpy()), t
        assert torch.equal(vb, v8)
        if t % 6 == 5:                                                                           # the world model, both formats
            cap = n * e.kmax
            pf8 = torch.empty((cap, 2, H, W), dtype=torch.uint8, device=dev); pfb = torch.zeros((cap, 2, SW), dtype=torch.int32, device=dev)
            pv8 = torch.empty((cap, 2, 19), dtype=torch.uint8, device=dev); pvb = torch.empty((cap, 2, 19), dtype=torch.uint8, device=dev)
            e.expand_dev(pl, K, offs, cap, field=pf8, vec=pv8, perspective=persp)
            e.expand_dev(pl, K, offs, cap, field=pfb, vec=pvb, perspective=persp)
            torch.cuda.synchronize()
            tot = int(offs[-1].item())
            assert np.array_equal(unpack(pfb[:tot]), pf8[:tot].cpu().numpy()) and torch.equal(pv8[:tot], pvb[:tot])
    e.close()
