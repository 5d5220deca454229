"""CPU: the N>1 sharding logic (tortoise_tts_b200/parallel.py) with world_size 2 over gloo."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tortoise_tts_b200 import parallel


def test_shard_range_partitions():
    for total in (1, 7, 16, 96, 256):
        for ws in (1, 2, 4, 8):
            spans = [parallel.shard_range(total, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
    assert parallel.shard_range(256, 3, 8) == (96, 128)
    assert [parallel.owner_of(j, 4) for j in range(6)] == [0, 1, 2, 3, 0, 1]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, total, L):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        g = torch.Generator().manual_seed(0)
        all_scores = torch.randn(total, generator=g)
        all_codes = torch.randint(0, 8194, (total, L), generator=g, dtype=torch.int32)
        lo, hi = parallel.shard_range(total, rank, ws)
        s, c = parallel.gather_candidates(all_scores[lo:hi].clone(), all_codes[lo:hi].clone(), total)
        assert torch.equal(s, all_scores) and torch.equal(c, all_codes)        # bit-exact, global order
        best = torch.topk(s, k=3).indices
        assert torch.equal(best, torch.topk(all_scores, k=3).indices)           # same ranking on every rank
        for j in range(3):
            owner = parallel.owner_of(j, ws)
            wav = torch.full((100 + j,), float(j + 1)) if rank == owner else None
            got = parallel.broadcast_from_owner(wav, 100 + j, owner, torch.device("cpu"))
            assert got.shape == (100 + j,) and bool((got == j + 1).all())
    finally:
        dist.destroy_process_group()


def test_gather_and_broadcast_world2():
    port = _free_port()
    mp.spawn(_worker, args=(2, port, 13, 9), nprocs=2, join=True)


def test_render_plan():
    assert parallel.render_plan(0, 1, True) == (0, None)
    assert parallel.render_plan(3, 4, False) == (3, None)
    assert [parallel.render_plan(j, 8, True) for j in range(5)] == [(0, 0), (2, 1), (4, 2), (6, 3), (0, 0)]
    assert parallel.render_plan(0, 2, True) == (0, 0)


def _pair_worker(rank, ws, port, ret):
    """CFG branches split over a rank pair (one all-gather per step) == both branches on one rank."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import lib_emu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lib_emu.install()
        from tortoise_tts_b200.config import ModelConfig
        from tortoise_tts_b200.synth import synth_all
        from tortoise_tts_b200.diffusion_engine import DiffusionEngine
        cfg = ModelConfig.small()
        sds = synth_all(cfg, seed=0, suppress_stop=False)
        g = torch.Generator().manual_seed(5)
        N, iters = 8, 3
        S = N * 4 * 24000 // 22050
        lat = torch.randn(N, cfg.ar_dim, generator=g)
        cond = torch.randn(2 * cfg.diff_dim, generator=g) * 0.3
        n0 = torch.randn(100, S, generator=g)
        sn = torch.randn(iters, 100, S, generator=g)
        groups, npairs = parallel.pair_groups()
        assert npairs == 1
        eng = DiffusionEngine(sds["diffusion"], cfg, device="cpu")
        mel_pair = eng.sample(lat, cond, iters, n0, sn, cond_free=True, cond_free_k=2.0, use_graph=False,
                              pair=(groups[0], rank))
        eng2 = DiffusionEngine(sds["diffusion"], cfg, device="cpu")
        mel_one = eng2.sample(lat, cond, iters, n0, sn, cond_free=True, cond_free_k=2.0, use_graph=False)
        assert (mel_pair - mel_one).abs().max().item() < 1e-4
    finally:
        dist.destroy_process_group()


def test_cfg_pair_split_world2():
    port = _free_port()
    mp.spawn(_pair_worker, args=(2, port, None), nprocs=2, join=True)


def test_gather_and_broadcast_world4_ragged():
    """13 candidates over 4 ranks (4, 4, 4, 1): the padded all-gather still returns the global order bit-exactly."""
    port = _free_port()
    mp.spawn(_worker, args=(4, port, 13, 9), nprocs=4, join=True)


def _pair_worker4(rank, ws, port, ret):
    """world 4, k = 1: the pair (0, 1) renders the only selected candidate, ranks 2 and 3 take no part in the denoiser
    but must create the pair groups and join the final broadcast (the layout bench.py --gpus 4/8 runs)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import lib_emu
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        lib_emu.install()
        from tortoise_tts_b200.config import ModelConfig
        from tortoise_tts_b200.synth import synth_all
        from tortoise_tts_b200.diffusion_engine import DiffusionEngine
        cfg = ModelConfig.small()
        sds = synth_all(cfg, seed=0, suppress_stop=False)
        g = torch.Generator().manual_seed(5)
        N, iters = 8, 2
        S = N * 4 * 24000 // 22050
        lat = torch.randn(N, cfg.ar_dim, generator=g)
        cond = torch.randn(2 * cfg.diff_dim, generator=g) * 0.3
        n0 = torch.randn(100, S, generator=g)
        sn = torch.randn(iters, 100, S, generator=g)
        groups, npairs = parallel.pair_groups()
        assert npairs == 2
        owner, p = parallel.render_plan(0, ws, True)
        assert (owner, p) == (0, 0)
        mel = None
        if rank in (owner, owner + 1):
            eng = DiffusionEngine(sds["diffusion"], cfg, device="cpu")
            mel = eng.sample(lat, cond, iters, n0, sn, cond_free=True, cond_free_k=2.0, use_graph=False,
                             pair=(groups[p], rank - owner))
        n = torch.tensor([mel.numel() if rank == owner else 0], dtype=torch.int64)
        dist.broadcast(n, src=owner)
        got = parallel.broadcast_from_owner(mel.reshape(-1) if rank == owner else None, int(n.item()), owner,
                                            torch.device("cpu"))
        assert got.numel() == 100 * S and bool(torch.isfinite(got).all())
        ref = DiffusionEngine(sds["diffusion"], cfg, device="cpu").sample(lat, cond, iters, n0, sn, cond_free=True,
                                                                          cond_free_k=2.0, use_graph=False)
        assert (got.reshape(100, S) - ref).abs().max().item() < 1e-4
    finally:
        dist.destroy_process_group()


def test_cfg_pair_split_world4():
    port = _free_port()
    mp.spawn(_pair_worker4, args=(4, port, None), nprocs=4, join=True)


def _utt_worker(rank, ws, port, n_utt):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from tortoise_tts_b200.text import utterance_plan
        plan = utterance_plan(n_utt, ws)
        with parallel.single_rank():
            assert parallel.world() == (0, 1)          # inside: the rank works alone, no collective is issued
            parts = {u: torch.full((50 + 7 * u,), float(u + 1)) for u in range(n_utt) if plan[u] == rank}
        assert parallel.world() == (rank, ws)
        got = parallel.exchange_utterances(parts, plan, torch.device("cpu"))
        assert len(got) == n_utt
        for u, w in enumerate(got):
            assert w.shape == (50 + 7 * u,) and bool((w == u + 1).all())
    finally:
        dist.destroy_process_group()


def test_utterance_sharding_world3():
    """Long-form mode: utterance u on rank u % G, ragged lengths, every rank gets all of them."""
    port = _free_port()
    mp.spawn(_utt_worker, args=(3, port, 7), nprocs=3, join=True)
