"""A small but complete tour of the C ABI on three field sizes (reset, keyed / placement / packed steps with fused reset, observe
in every dtype, enumerate, expand in both finalize modes, select, snapshot, get/set/copy game, tile planes): every call must return
B2T_OK and leave the device healthy (the capacity counters stay zero).  compute-sanitizer is closed on the GPU pool, so the bounds
evidence for the kernels is tests/hostsim (the same headers compiled with B2T_BOUNDS_CHECK on the CPU) plus this tour."""
import numpy as np
import pytest

import harness


@pytest.mark.gpu
def test_c_abi_tour():
    import torch
    from b200tetris import Engine, host_array
    dev = torch.device("cuda:0")
    for H, W, pieces, n in ((22, 10, (0, 1, 2, 3, 4, 5, 6), 333), (16, 8, (0, 6), 70), (40, 20, (0, 1, 2, 3, 4, 5, 6), 45)):
        e = Engine(n, H, W, pieces, piece_stream_len=256, hole_stream_len=64, tiles=(W == 10))
        gids = np.arange(n)
        e.reset(harness.seeds_for(gids, np.zeros(n, int)))
        pol = harness.Policy("junk", lmax=16)
        K = torch.empty(n, dtype=torch.int32, device=dev); offs = torch.empty(n + 1, dtype=torch.int32, device=dev)
        cap = n * e.kmax
        for t in range(14):
            player = ((gids + t) & 1).astype(np.uint8)
            keys, klen, _ = pol.choose(None, t, gids, player)
            e.step(player, keys, klen, want_obs=True, want_scalars=True)
            e.step(player, keys, None, auto_reset=True)
            wk = host_array(keys.shape, np.uint8, write_combined=True)      # b2t_host_alloc: pinned, write-combined inputs
            wk[...] = keys
            wr = host_array(n, np.int8)
            import ctypes as C
            from b200tetris.engine import _hp
            assert e.lib.b2t_step_host(e._h, None, n, _hp(player, C.c_uint8), _hp(wk, C.c_uint8), None, keys.shape[1], None, 1, 0, None,
                                       _hp(wr, C.c_int8), None, None, None, None, None, None, 0, 1) == 0
            pl = torch.from_numpy(player).to(dev)
            for dt in (torch.uint8, torch.float16, torch.float32):
                f = torch.empty((n, 2, H, W), dtype=dt, device=dev); v = torch.empty((n, 2, 19), dtype=dt, device=dev)
                e.observe_dev(field=f, vec=v, me=pl)
            for fin in (True, False):
                pf = torch.empty((cap, 2, H, W), dtype=torch.uint8, device=dev); pv = torch.empty((cap, 2, 19), dtype=torch.uint8, device=dev)
                pk = torch.empty((cap, e.lmax), dtype=torch.uint8, device=dev); pr = torch.empty(cap, dtype=torch.int8, device=dev)
                e.expand_dev(pl, K, offs, cap, finalize=fin, keys=pk, reward=pr, field=pf, vec=pv, perspective=True)
            scores = torch.rand(cap, device=dev)
            sel = torch.empty(n, dtype=torch.int32, device=dev)
            e.select_dev(scores, offs, sel)
            seeds = torch.from_numpy(harness.seeds_for(gids, np.full(n, t + 1)).astype(np.int64)).to(dev).to(torch.int32)
            e.step_dev(pl, placement=sel, reset_seeds=seeds, done=torch.empty(n, dtype=torch.uint8, device=dev))
            packed = np.ascontiguousarray(keys[:, 0::2] | (keys[:, 1::2] << 4))
            e.step_dev(pl, torch.from_numpy(packed).to(dev), None, packed_keys=True, auto_reset=True)
        e.snapshot(0); e.restore(0)
        raw, fld = e.get_game(3); e.set_game(4, raw, fld); e.copy_game(5, e, 3)
        e.enumerate(np.zeros(n, np.uint8))
        torch.cuda.synchronize()
        st = e.stats()
        assert st["capacity_overflow"] == 0 and st["garbage_queue_overflow"] == 0, st
        e.close()