"""Generates tests/golden/small_v1.pt from the UNMODIFIED reference modules (run in the build container only:
`python tests/golden/make_golden.py`). The fixture pins the oracle (CPU tests) and the CUDA path (GPU tests) to
outputs of the reference's own PyTorch code on the synthetic small checkpoint (ModelConfig.small(), seed 0).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from tortoise_tts_b200.config import ModelConfig  # noqa: E402
from tortoise_tts_b200.synth import synth_all  # noqa: E402
from oracle.ref_build import build_reference_models  # noqa: E402

TEXT = [42, 2, 194, 91, 24, 2, 243, 190, 2, 182, 37, 2, 0]  # 12 BPE ids + the api.py:391 pad


def main():
    torch.manual_seed(1234)
    cfg = ModelConfig.small()
    sds = synth_all(cfg, seed=0, suppress_stop=False)
    m = build_reference_models(cfg, sds, kv_cache=True)
    uv, dm, clvp, voc = m["autoregressive"], m["diffusion"], m["clvp"], m["vocoder"]
    out = {"text": torch.tensor(TEXT), "weights_checksum": {k: float(sum(v.double().sum() for v in sd.values()))
                                                           for k, sd in sds.items()}}
    with torch.no_grad():
        # ---- AR: cached-path logits (reference kv-cache position rule) and recompute-path logits
        cond = torch.randn(1, cfg.ar_dim)
        codes = torch.randint(0, 8192, (2, 6))
        out["ar_cond"], out["ar_codes"] = cond, codes
        inf = uv.inference_model
        text = torch.tensor(TEXT).unsqueeze(0)
        ti = torch.nn.functional.pad(text, (0, 1), value=0)
        ti = torch.nn.functional.pad(ti, (1, 0), value=cfg.start_text_token)
        emb = uv.text_embedding(ti) + uv.text_pos_embedding(ti)
        emb = torch.cat([cond.unsqueeze(1), emb], dim=1)
        inf.store_mel_emb(emb)
        fake = torch.full((2, emb.shape[1] + 1), 1, dtype=torch.long)
        fake[:, -1] = cfg.start_mel_token
        ids, past, got = fake, None, []
        for j in range(codes.shape[1] + 1):
            am = torch.ones_like(ids)
            o = inf(input_ids=ids if past is None else ids[:, -1:], past_key_values=past, attention_mask=am,
                    use_cache=True, return_dict=True)
            past = o.past_key_values
            got.append(o.logits[:, -1])
            if j < codes.shape[1]:
                ids = torch.cat([ids, codes[:, j:j + 1]], dim=1)
        out["ar_logits_kv"] = torch.stack(got, dim=1)
        inf.kv_cache = False
        ids = torch.cat([fake, codes], dim=1)
        out["ar_logits_recompute"] = inf(input_ids=ids, attention_mask=torch.ones_like(ids),
                                         return_dict=True).logits[:, emb.shape[1]:]
        inf.kv_cache = True
        # ---- latents
        lcodes = torch.randint(0, 8192, (2, 10))
        out["lat_codes"] = lcodes
        out["latents"] = uv(cond.repeat(2, 1), text.repeat(2, 1), torch.tensor([text.shape[-1]]), lcodes,
                            torch.tensor([lcodes.shape[-1] * uv.mel_length_compression]), return_latent=True,
                            clip_inputs=False)
        # ---- CLVP
        ccodes = torch.randint(0, 8192, (3, 24))
        out["clvp_codes"] = ccodes
        out["clvp_scores"] = clvp(text.repeat(3, 1), ccodes, return_loss=False)
        # ---- diffusion
        from tortoise.api import load_discrete_vocoder_diffuser, do_spectrogram_diffusion
        N = 10
        lat = torch.randn(1, N, cfg.ar_dim)
        dcond = torch.randn(1, 2 * cfg.diff_dim)
        S = N * 4 * 24000 // 22050
        out["diff_latents"], out["diff_cond"] = lat, dcond
        out["code_emb"] = dm.timestep_independent(lat, dcond, S, False)
        x = torch.randn(1, 100, S)
        out["diff_x"] = x
        out["diff_fwd_cond"] = dm(x, torch.tensor([3979]), precomputed_aligned_embeddings=out["code_emb"])
        out["diff_fwd_uncond"] = dm(x, torch.tensor([3979]), precomputed_aligned_embeddings=out["code_emb"],
                                    conditioning_free=True)
        iters = 6
        diffuser = load_discrete_vocoder_diffuser(desired_diffusion_steps=iters, cond_free=True, cond_free_k=2.0)
        torch.manual_seed(77)
        out["diff_mel"] = do_spectrogram_diffusion(dm, diffuser, lat, dcond, temperature=1.0, verbose=False)
        torch.manual_seed(77)
        out["diff_noise0"] = torch.randn(1, 100, S)
        out["diff_step_noise"] = torch.stack([torch.randn(1, 100, S) for _ in range(iters)])
        out["diff_iters"] = iters
        # ---- vocoder
        mel = torch.randn(1, 100, 14) * 2 - 5
        z = torch.randn(1, 64, 24)
        out["voc_mel"], out["voc_z"] = mel, z
        out["voc_wav"] = voc.inference(mel, z)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "small_v1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
