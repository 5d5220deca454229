"""TEST INFRASTRUCTURE ONLY — build the reference's own nn.Modules (from /root/reference)
for a ModelConfig and load a synthetic checkpoint into them.  Used to validate oracle/*.py
and to generate tests/golden/*.  Mirrors the constructor calls in tortoise/api.py:217-238.
"""
import torch

from .ref_shims import load_reference


def build_reference_models(cfg, sds, kv_cache=True):
    load_reference()
    from tortoise.models.autoregressive import UnifiedVoice
    from tortoise.models.diffusion_decoder import DiffusionTts
    from tortoise.models.clvp import CLVP
    from tortoise.models.vocoder import UnivNetGenerator

    ar = UnifiedVoice(max_mel_tokens=cfg.max_mel_tokens, max_text_tokens=cfg.max_text_tokens,
                      max_conditioning_inputs=cfg.max_conditioning_inputs, layers=cfg.ar_layers,
                      model_dim=cfg.ar_dim, heads=cfg.ar_heads, number_text_tokens=cfg.number_text_tokens,
                      start_text_token=cfg.start_text_token, checkpointing=False,
                      train_solo_embeddings=False).cpu().eval()
    # the reference ConditioningEncoder has a fixed 6 blocks; drop missing ones for reduced configs
    missing = ar.load_state_dict(sds["autoregressive"], strict=(cfg.cond_enc_blocks == 6))
    ar.post_init_gpt2_config(use_deepspeed=False, kv_cache=kv_cache, half=False)

    diff = DiffusionTts(model_channels=cfg.diff_dim, num_layers=cfg.diff_layers, in_channels=cfg.diff_in_channels,
                        out_channels=cfg.diff_out_channels, in_latent_channels=cfg.ar_dim,
                        in_tokens=cfg.diff_in_tokens, dropout=0, use_fp16=False, num_heads=cfg.diff_heads,
                        layer_drop=0, unconditioned_percentage=0).cpu().eval()
    diff.load_state_dict(sds["diffusion"], strict=True)

    clvp = CLVP(dim_text=cfg.clvp_dim, dim_speech=cfg.clvp_dim, dim_latent=cfg.clvp_dim,
                num_text_tokens=cfg.clvp_text_tokens, text_enc_depth=cfg.clvp_depth, text_seq_len=350,
                text_heads=cfg.clvp_heads, num_speech_tokens=cfg.clvp_speech_tokens,
                speech_enc_depth=cfg.clvp_depth, speech_heads=cfg.clvp_heads, speech_seq_len=430,
                use_xformers=True).cpu().eval()
    clvp.load_state_dict(sds["clvp"], strict=True)

    voc = UnivNetGenerator().cpu()
    voc.load_state_dict(sds["vocoder"], strict=True)
    voc.eval(inference=True)
    from tortoise.models.cvvp import CVVP
    cvvp = CVVP(model_dim=cfg.cvvp_dim, transformer_heads=cfg.cvvp_heads, dropout=0, mel_codes=cfg.clvp_speech_tokens,
                conditioning_enc_depth=cfg.cvvp_depth, cond_mask_percentage=0, speech_enc_depth=cfg.cvvp_depth,
                speech_mask_percentage=0, latent_multiplier=1).eval()          # api.py:254-255
    cvvp.load_state_dict(sds["cvvp"], strict=True)
    from tortoise.models.hifigan_decoder import HifiganGenerator
    hifi = HifiganGenerator(in_channels=cfg.ar_dim, out_channels=1, resblock_type="1",
                            resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]], resblock_kernel_sizes=[3, 7, 11],
                            upsample_kernel_sizes=[16, 16, 4, 4], upsample_initial_channel=cfg.hifi_channels,
                            upsample_factors=[8, 8, 2, 2], cond_channels=cfg.ar_dim).eval()      # api_fast.py:221-224
    hifi.device = torch.device("cpu")
    hifi.load_state_dict(sds["hifigan"], strict=True)
    return {"autoregressive": ar, "diffusion": diff, "clvp": clvp, "vocoder": voc, "cvvp": cvvp, "hifigan": hifi,
            "missing": missing}
