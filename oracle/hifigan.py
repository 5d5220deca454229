"""TEST INFRASTRUCTURE ONLY (CPU oracle) — restatement of HifiganGenerator.inference (the decoder of the reference's
`api_fast` path: GPT latents -> 24 kHz waveform).

Plain torch-fp32 functional code over the reference `hifidecoder.pth` state_dict (weight norm in place). Follows
tortoise/models/hifigan_decoder.py: `inference` 270-294 (two linear interpolations of the latents: x 4, then x 24000/22050;
the speaker latent through `cond_layer`), `forward` 240-268 (conv_pre + cond -> 4 x [leaky_relu 0.1 -> ConvTranspose1d ->
mean of three ResBlock1] -> leaky_relu (default slope 0.01) -> conv_post -> tanh), `ResBlock1.forward` 83-97.
Pinned against the reference module by tests/test_oracle_vs_reference.py::test_hifigan.
"""
import torch
import torch.nn.functional as F

UP_FACTORS = (8, 8, 2, 2)
RES_KERNELS = (3, 7, 11)
RES_DILATIONS = (1, 3, 5)
LRELU_SLOPE = 0.1


def _w(sd, p):
    """weight_norm(dim=0): w = g * v / ||v|| with the norm over every dim but 0 (torch.nn.utils.weight_norm)."""
    v, g = sd[p + "weight_v"], sd[p + "weight_g"]
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(g.shape)
    return v * (g / n)


def upsample_latents(latents):
    """hifigan_decoder.py:283-292. latents [B, L, C] -> [B, C, T]."""
    up = F.interpolate(latents.transpose(1, 2), scale_factor=[1024 / 256], mode="linear")
    return F.interpolate(up, scale_factor=[24000 / 22050], mode="linear")


def resblock1(sd, p, x, k):
    for m, d in enumerate(RES_DILATIONS):
        xt = F.leaky_relu(x, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, p + f"convs1.{m}."), sd[p + f"convs1.{m}.bias"], dilation=d, padding=(k * d - d) // 2)
        xt = F.leaky_relu(xt, LRELU_SLOPE)
        xt = F.conv1d(xt, _w(sd, p + f"convs2.{m}."), sd[p + f"convs2.{m}.bias"], padding=(k - 1) // 2)
        x = xt + x
    return x


def forward(sd, x, g):
    """x [B, C_in, T], g [B, C_in, 1] -> [B, 1, 256 T]."""
    o = F.conv1d(x, _w(sd, "conv_pre."), sd["conv_pre.bias"], padding=3)
    o = o + F.conv1d(g, sd["cond_layer.weight"], sd["cond_layer.bias"])
    for i, u in enumerate(UP_FACTORS):
        o = F.leaky_relu(o, LRELU_SLOPE)
        o = F.conv_transpose1d(o, _w(sd, f"ups.{i}."), sd[f"ups.{i}.bias"], stride=u, padding=u // 2)
        z = None
        for j, k in enumerate(RES_KERNELS):
            r = resblock1(sd, f"resblocks.{i * len(RES_KERNELS) + j}.", o, k)
            z = r if z is None else z + r
        o = z / len(RES_KERNELS)
    o = F.leaky_relu(o)
    o = F.conv1d(o, _w(sd, "conv_post."), sd["conv_post.bias"], padding=3)
    return torch.tanh(o)


def inference(sd, latents, speaker):
    """latents [1, L, C] (GPT latents), speaker [1, C] (auto conditioning latent) -> waveform [1, 1, 256 T]."""
    return forward(sd, upsample_latents(latents), speaker.unsqueeze(0).transpose(1, 2))
