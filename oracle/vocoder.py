"""TEST INFRASTRUCTURE ONLY (CPU oracle) — restatement of UnivNetGenerator.inference.

Plain torch-fp32 functional code over the reference `vocoder.pth['model_g']` state_dict in
its weight-norm form. Follows models/vocoder.py:
  * remove_weight_norm (w = g * v / ||v||, norm over all dims but 0)        290-298
  * inference: append 10 frames of -11.5129, z ~ N(0,1)[64, S+10], crop, clamp 300-312
  * forward: conv_pre k7 reflect -> 3 LVCBlocks -> LeakyReLU, conv_post k7 reflect, tanh  267-282
  * LVCBlock.forward: LeakyReLU, ConvTranspose1d(2s, stride s, pad s/2+s%2, out_pad s%2);
    per layer: LeakyReLU, dilated conv3, LeakyReLU, LVC, sigmoid*tanh gate    155-180
  * KernelPredictor.forward                                                   66-93
  * location_variable_convolution, closed form:
      o[oc, f*hop+s] = sum_{i,k} ypad[i, f*hop+s+k] * K[i, oc, k, f] + Bias[oc, f]   182-216
"""
import torch
import torch.nn.functional as F

from tortoise_tts_b200.config import VOC_STRIDES, VOC_DILATIONS, VOC_LRELU


def fold_weight_norm(sd, prefix):
    g, v = sd[prefix + "weight_g"], sd[prefix + "weight_v"]
    norm = v.reshape(v.shape[0], -1).norm(dim=1).reshape(v.shape[0], *([1] * (v.dim() - 1)))
    return g * v / norm


def kernel_predictor(sd, p, c):
    """c [1,100,F] -> kernels [4, 32, 64, 3, F], bias [4, 64, F]."""
    w = lambda n: fold_weight_norm(sd, p + n)
    b = lambda n: sd[p + n + "bias"]
    c = F.leaky_relu(F.conv1d(c, w("input_conv.0."), b("input_conv.0."), padding=2), VOC_LRELU)
    for r in range(3):
        h = F.leaky_relu(F.conv1d(c, w(f"residual_convs.{r}.1."), b(f"residual_convs.{r}.1."), padding=1), VOC_LRELU)
        h = F.leaky_relu(F.conv1d(h, w(f"residual_convs.{r}.3."), b(f"residual_convs.{r}.3."), padding=1), VOC_LRELU)
        c = c + h
    k = F.conv1d(c, w("kernel_conv."), b("kernel_conv."), padding=1)
    bb = F.conv1d(c, w("bias_conv."), b("bias_conv."), padding=1)
    Fr = c.shape[-1]
    nl = len(VOC_DILATIONS)
    ch = k.shape[1] // (nl * 3 * 2)
    ch = int(round((ch) ** 0.5))  # k channels = nl * ch * 2ch * 3
    return k.view(nl, ch, 2 * ch, 3, Fr), bb.view(nl, 2 * ch, Fr)


def lvc(y, K, Bias, hop):
    """y [C, L] (L = F*hop), K [C, 2C, 3, F], Bias [2C, F] -> [2C, L]."""
    C, L = y.shape
    Fr = K.shape[-1]
    ypad = F.pad(y, (1, 1))
    # windows [C, F, hop+2]
    win = ypad.unfold(1, hop + 2, hop)
    taps = win.unfold(2, 3, 1)  # [C, F, hop, 3]
    o = torch.einsum("ifsk,iokf->ofs", taps, K)
    o = o + Bias[:, :, None]
    return o.reshape(K.shape[1], L)


def lvc_block(sd, b, x, c, stride, hop):
    p = f"res_stack.{b}."
    x = F.leaky_relu(x, VOC_LRELU)
    x = F.conv_transpose1d(x, fold_weight_norm(sd, p + "convt_pre.1."), sd[p + "convt_pre.1.bias"],
                           stride=stride, padding=stride // 2 + stride % 2, output_padding=stride % 2)
    K, Bias = kernel_predictor(sd, p + "kernel_predictor.", c)
    ch = x.shape[1]
    for i, d in enumerate(VOC_DILATIONS):
        y = F.leaky_relu(x, VOC_LRELU)
        y = F.conv1d(y, fold_weight_norm(sd, p + f"conv_blocks.{i}.1."), sd[p + f"conv_blocks.{i}.1.bias"],
                     padding=d, dilation=d)
        y = F.leaky_relu(y, VOC_LRELU)
        o = lvc(y[0], K[i], Bias[i], hop).unsqueeze(0)
        x = x + torch.sigmoid(o[:, :ch]) * torch.tanh(o[:, ch:])
    return x


def inference(sd, mel, z):
    """mel [1,100,S], z [1,64,S+10] -> audio [1,1,256*S] in [-1,1] (vocoder.py:300-312)."""
    zero = torch.full((mel.shape[0], mel.shape[1], 10), -11.5129)
    c = torch.cat((mel, zero), dim=2)
    x = F.conv1d(F.pad(z, (3, 3), mode="reflect"), fold_weight_norm(sd, "conv_pre."), sd["conv_pre.bias"])
    hop = 1
    for b, s in enumerate(VOC_STRIDES):
        hop *= s
        x = lvc_block(sd, b, x, c, s, hop)
    x = F.leaky_relu(x, VOC_LRELU)
    x = torch.tanh(F.conv1d(F.pad(x, (3, 3), mode="reflect"), fold_weight_norm(sd, "conv_post.1."),
                            sd["conv_post.1.bias"]))
    x = x[:, :, :-(256 * 10)]
    return x.clamp(min=-1, max=1)
