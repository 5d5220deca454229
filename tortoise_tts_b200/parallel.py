"""Candidate / utterance sharding across the GPUs of one box.

One process per GPU (torch.distributed, NCCL over NVLink; gloo in the CPU tests). The hot path shards without any
per-step collective: autoregressive candidates are independent given (text tokens, conditioning latent), so rank r
decodes candidates [r*B/G, (r+1)*B/G) and scores them with CLVP locally; ONE all-gather of {scores f32, codes i32}
lets every rank run the same top-k; the k selected candidates' latents / diffusion / vocoder run on rank j % G and
the waveforms are broadcast from their owners. The reference has no distributed code: this is new.
"""
import torch
import torch.distributed as dist


_SINGLE = 0   # > 0 inside single_rank(): this rank works alone (utterance-level sharding)


def world():
    if _SINGLE == 0 and dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


class single_rank:
    """Context in which `world()` reports one rank: the calls inside do all their work locally and issue no collective.
    Used when whole utterances, not candidates, are spread over the GPUs."""

    def __enter__(self):
        global _SINGLE
        _SINGLE += 1
        return self

    def __exit__(self, *exc):
        global _SINGLE
        _SINGLE -= 1
        return False


def broadcast_seed(seed, device):
    """Rank 0's seed on every rank (one int64 broadcast); the identity without a process group."""
    rank, ws = world()
    if ws == 1:
        return int(seed)
    backend = dist.get_backend()
    t = torch.tensor([int(seed)], dtype=torch.int64, device=device if backend == "nccl" else "cpu")
    dist.broadcast(t, src=0)
    return int(t.item())


def shard_range(total, rank, world_size):
    """Contiguous, balanced split of `total` candidates: returns (lo, hi) for `rank`."""
    per = (total + world_size - 1) // world_size
    lo = min(total, rank * per)
    return lo, min(total, lo + per)


def owner_of(j, world_size):
    """Rank that renders the j-th selected candidate (diffusion + vocoder)."""
    return j % world_size


_PAIR_GROUPS = None


def pair_groups():
    """Process groups of rank pairs (2p, 2p+1) used to split the two classifier-free-guidance branches of the denoiser
    over two GPUs. Created collectively (every rank calls new_group for every pair, in the same order) on first use.
    Returns (groups list, number of pairs) or (None, 0) when the world has fewer than 2 ranks."""
    global _PAIR_GROUPS
    rank, ws = world()
    if ws < 2:
        return None, 0
    if _PAIR_GROUPS is None:
        _PAIR_GROUPS = [dist.new_group([2 * p, 2 * p + 1]) for p in range(ws // 2)]
    return _PAIR_GROUPS, ws // 2


def render_plan(j, world_size, cond_free):
    """Who renders (diffusion + vocoder) the j-th selected candidate: returns (owner rank, pair index or None).
    With CFG and >= 2 ranks, pair p = j % (G//2) = ranks (2p, 2p+1) share the denoiser; rank 2p owns the waveform."""
    if cond_free and world_size >= 2:
        p = j % (world_size // 2)
        return 2 * p, p
    return owner_of(j, world_size), None


def gather_candidates(scores, codes, total):
    """All-gather of per-rank CLVP scores [b] and codes [b, L] -> ([total], [total, L]) in global candidate order.
    One fused collective: scores are bit-cast into an extra int32 column of the codes buffer."""
    rank, ws = world()
    if ws == 1:
        return scores, codes
    per = (total + ws - 1) // ws
    L = codes.shape[1]
    buf = torch.zeros(per, L + 1, dtype=torch.int32, device=codes.device)
    n = codes.shape[0]
    buf[:n, :L] = codes
    buf[:n, L] = scores.contiguous().view(torch.int32)
    out = torch.empty(ws * per, L + 1, dtype=torch.int32, device=codes.device)
    dist.all_gather_into_tensor(out, buf) if codes.is_cuda else dist.all_gather(list(out.chunk(ws)), buf)
    keep = torch.cat([out[r * per: r * per + (shard_range(total, r, ws)[1] - shard_range(total, r, ws)[0])]
                      for r in range(ws)], dim=0)
    return keep[:, L].contiguous().view(torch.float32), keep[:, :L].contiguous()


def broadcast_from_owner(t, numel, owner, device, dtype=torch.float32):
    """Every rank ends up with the owner's 1-D tensor of `numel` elements."""
    rank, ws = world()
    if ws == 1:
        return t
    buf = t.contiguous() if rank == owner else torch.empty(numel, dtype=dtype, device=device)
    dist.broadcast(buf, src=owner)
    return buf


def exchange_utterances(parts, plan, device, dtype=torch.float32):
    """`parts[u]` = 1-D waveform of utterance u on its owner `plan[u]` (absent elsewhere). Returns the list of all
    utterances on every rank: per utterance one length broadcast and one payload broadcast from its owner."""
    rank, ws = world()
    out = []
    for u, owner in enumerate(plan):
        if ws == 1:
            out.append(parts[u])
            continue
        mine = parts[u].to(device=device, dtype=dtype).reshape(-1) if owner == rank else None
        n = torch.tensor([mine.numel() if owner == rank else 0], dtype=torch.int64, device=device)
        dist.broadcast(n, src=owner)
        out.append(broadcast_from_owner(mine, int(n.item()), owner, device, dtype))
    return out


class PairExchange:
    """Exchange area of a CFG rank pair over peer memory (CUDA IPC + NVLink stores), used by `lib.pair_exchange`
    (csrc/misc.cu) inside the captured denoiser step. Layout of `area`: fp32 [2 step parities][2 branches][n]; `flags`:
    int32, [0:2] written by the partner. Both live in their own cudaMalloc allocations (lib.peer_alloc) and are mapped into
    the partner process with its device current. `ok` is False when the mapping could not be set up on BOTH ranks (the
    caller then keeps the NCCL all-gather)."""

    def __init__(self, group, my_idx, n, device):
        import os
        from . import lib
        self.n, self.my_idx = int(n), int(my_idx)
        self.epoch = torch.zeros(1, dtype=torch.int32, device=device)
        self.done = torch.zeros(1, dtype=torch.int32, device=device)
        self.err = torch.zeros(1, dtype=torch.int32, device=device)
        self.area = self.flags = self.peer_area = self.peer_flags = None
        ok = os.environ.get("TTB_PEER_EXCHANGE", "1") == "1" and torch.device(device).type == "cuda"
        mine = None
        if ok:
            try:
                self.area, ha = lib.peer_alloc(4 * 2 * 2 * self.n, torch.float32)
                self.flags, hf = lib.peer_alloc(4 * 32, torch.int32)
                mine = (ha, hf)
            except lib.TtbError:
                ok = False
        metas = [None, None]
        dist.all_gather_object(metas, mine, group=group)
        peer = metas[1 - self.my_idx]
        if ok and peer is not None:
            try:
                self.peer_area = lib.peer_open(peer[0], torch.float32, 2 * 2 * self.n)
                self.peer_flags = lib.peer_open(peer[1], torch.int32, 32)
            except lib.TtbError:
                ok = False
        else:
            ok = False
        oks = [None, None]
        dist.all_gather_object(oks, bool(ok), group=group)
        self.ok = all(bool(v) for v in oks)

    def new_sample(self):
        """Both ranks call this once per sampling loop, before its first step: flags of earlier loops become stale."""
        self.epoch.add_(4096)
