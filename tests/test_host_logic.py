"""CPU: host-side logic of the engine (schedule tables, relative-position tables, weight regrouping)."""
import numpy as np
import torch

from tortoise_tts_b200.config import ModelConfig
from tortoise_tts_b200 import diffusion_engine as de


def test_schedule_matches_oracle():
    from oracle import diffusion as od
    for iters in (5, 30, 80, 200, 400):
        tmap, tables = de.make_schedule(iters)
        s = od.make_schedule(iters)
        assert list(tmap) == list(s["timestep_map"])
        names = ["sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped",
                 "log_betas", "posterior_mean_coef1", "posterior_mean_coef2"]
        for r, n in enumerate(names):
            assert np.array_equal(tables[r], s[n].astype(np.float32)), n


def test_rel_pos_table_matches_oracle():
    from oracle import diffusion as od
    torch.manual_seed(0)
    emb = torch.randn(32, 4)
    T = 37
    tab = de._rel_pos_table(emb, T, 8.0)          # [H, 2T-1]
    full = od.rel_pos_bias(emb, T, 8.0)           # [H, T, T]
    for i in (0, 5, 36):
        for j in (0, 17, 36):
            assert torch.allclose(tab[:, j - i + T - 1], full[:, i, j])


def test_rel_pos_table_saturates_at_max_distance():
    """|k - q| >= 64 -> constant bias per side: the property `bias_sat=64` promises to the flash kernel."""
    torch.manual_seed(1)
    T = 300
    tab = de._rel_pos_table(torch.randn(32, 3), T, 8.0)
    c = T - 1
    assert bool((tab[:, : c - 64 + 1] == tab[:, :1]).all())
    assert bool((tab[:, c + 64:] == tab[:, -1:]).all())
    assert not bool((tab[:, c - 40] == tab[:, 0]).all())


def test_groups_rule():
    from oracle import diffusion as od
    for C in (16, 64, 128, 1024, 2048):
        assert de._groups_for(C) == od.groups_for(C)


def test_qkv_regroup():
    """per-head [q|k|v] rows -> [q heads | k heads | v heads] (arch_util.py:60-63)."""
    C, H = 128, 2
    ch = C // H
    idx = torch.arange(3 * C).reshape(H, 3, ch).permute(1, 0, 2).reshape(-1)
    # new row (which=1 (k), head=1, d=5) must come from old row head*192 + which*64 + d
    assert int(idx[1 * C + 1 * ch + 5]) == 1 * 3 * ch + 1 * ch + 5


def test_presets_match_reference_values():
    from tortoise_tts_b200.api import PRESETS
    assert PRESETS["standard"] == {"num_autoregressive_samples": 256, "diffusion_iterations": 200}
    assert PRESETS["ultra_fast"]["cond_free"] is False
    assert PRESETS["high_quality"]["diffusion_iterations"] == 400


def test_synth_layout_param_counts():
    from tortoise_tts_b200.synth import synth_all
    sds = synth_all(ModelConfig.small())
    assert "gpt.h.1.mlp.c_proj.weight" in sds["autoregressive"]
    assert sds["vocoder"]["res_stack.0.kernel_predictor.kernel_conv.weight_v"].shape == (24576, 64, 3)


def test_cpu_baseline_thread_probe():
    """bench.py's CPU arm must size its thread pool from what the process may use, never from os.cpu_count()."""
    import os
    from oracle import cpu_baseline as cb
    n = cb.host_threads()
    assert 1 <= n <= 64
    if hasattr(os, "sched_getaffinity"):
        assert n <= len(os.sched_getaffinity(0))
    assert cb.host_threads(cap=2) <= 2


def test_build_stamp_is_content_hash():
    """The kernel library is rebuilt when a source changes, not when file times change (a copied tree keeps no times)."""
    import os
    from tortoise_tts_b200 import build as b
    h1 = b._source_hash()
    assert len(h1) == 64 and h1 == b._source_hash()
    if os.path.exists(b.STAMP) and os.path.exists(b.LIB) and open(b.STAMP).read().strip() != h1:
        import warnings
        warnings.warn("libttb.so is older than csrc/: __graft_entry__.build() will recompile it")
