"""Oracle: the UAPS multi-decoder U-Net forward as pure functions of a state_dict
(TEST INFRASTRUCTURE).  Line numbers: /root/reference/utilities/UAPS_unet.py.

The reference model (``UNet_UAPS``) is an nn.Module tree that draws its
randomness from global generators.  This restatement evaluates the same
layers straight from the reference's state_dict keys with every random draw
injected, so the CUDA path and the reference can be compared on identical
inputs.  ``oracle/make_golden.py`` checks it against the imported reference
module (randomness patched to the injected values) in the build container.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

from .perturb_ref import dropout_ref, feature_dropout_ref, feature_noise_ref

FT_CHNS = (16, 32, 64, 128, 256)            # :212
ENC_DROPOUT = (0.05, 0.1, 0.2, 0.3, 0.5)    # :214
ENC_PREFIX = ("encoder.in_conv", "encoder.down1.maxpool_conv.1", "encoder.down2.maxpool_conv.1",
              "encoder.down3.maxpool_conv.1", "encoder.down4.maxpool_conv.1")
DECODERS = ("main_decoder", "aux_decoder1", "aux_decoder2", "aux_decoder3")


def _bn_train(x, sd, key, stats_out):
    """nn.BatchNorm2d in training mode (:38,:42): batch statistics, eps 1e-5, momentum 0.1."""
    w, b = sd[key + ".weight"], sd[key + ".bias"]
    if stats_out is None:
        return F.batch_norm(x, None, None, w, b, True, 0.1, 1e-5)
    rm = stats_out.get(key + ".running_mean", sd[key + ".running_mean"]).clone()
    rv = stats_out.get(key + ".running_var", sd[key + ".running_var"]).clone()
    y = F.batch_norm(x, rm, rv, w, b, True, 0.1, 1e-5)
    stats_out[key + ".running_mean"], stats_out[key + ".running_var"] = rm, rv
    return y


def conv_block_ref(x, sd, prefix, p_drop, keep_mask, stats_out=None):
    """ConvBlock :31-47: conv3x3 -> BN -> LeakyReLU(0.01) -> Dropout(p) -> conv3x3 -> BN -> LeakyReLU."""
    cc = prefix + ".conv_conv"
    y = F.conv2d(x, sd[cc + ".0.weight"], sd[cc + ".0.bias"], padding=1)
    y = F.leaky_relu(_bn_train(y, sd, cc + ".1", stats_out), 0.01)
    if p_drop > 0.0:
        y = dropout_ref(y, keep_mask, p_drop)                              # nn.Dropout(p) :40
    y = F.conv2d(y, sd[cc + ".4.weight"], sd[cc + ".4.bias"], padding=1)
    return F.leaky_relu(_bn_train(y, sd, cc + ".5", stats_out), 0.01)


def encoder_ref(x, sd, enc_keep: Sequence[torch.Tensor], stats_out=None) -> List[torch.Tensor]:
    """Encoder.forward :110-116 (MaxPool2d(2) before every block but the first, :56)."""
    feats = []
    for lvl in range(5):
        if lvl > 0:
            x = F.max_pool2d(x, 2)
        x = conv_block_ref(x, sd, ENC_PREFIX[lvl], ENC_DROPOUT[lvl], enc_keep[lvl], stats_out)
        feats.append(x)
    return feats


def decoder_ref(feats: Sequence[torch.Tensor], sd, name: str, stats_out=None) -> torch.Tensor:
    """Decoder.forward :141-153; UpBlock.forward :81-86 with bilinear=True (the ctor default :69 wins)."""
    x = feats[4]
    for i, skip in zip((1, 2, 3, 4), (feats[3], feats[2], feats[1], feats[0])):
        up = f"{name}.up{i}"
        x = F.conv2d(x, sd[up + ".conv1x1.weight"], sd[up + ".conv1x1.bias"])           # :83
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)       # :74-75,84
        x = torch.cat([skip, x], dim=1)                                                 # :85
        x = conv_block_ref(x, sd, up + ".conv", 0.0, None, stats_out)                   # :86
    return F.conv2d(x, sd[name + ".out_conv.weight"], sd[name + ".out_conv.bias"], padding=1)  # :138-139,152


def unet_uaps_ref(x: torch.Tensor, sd: Dict[str, torch.Tensor], rand: Dict[str, list],
                  stats_out: Optional[dict] = None):
    """UNet_UAPS.forward :224-233.

    rand = {"enc_keep": 5 bool masks (encoder nn.Dropout :40),
            "noise":    5 tensors of shape feature.shape[1:] (FeatureNoise :178),
            "aux2_keep":5 bool masks (F.dropout p=0.5 :157),
            "u":        5 floats (FeatureDropout :165)}
    """
    feats = encoder_ref(x, sd, rand["enc_keep"], stats_out)
    main = decoder_ref(feats, sd, DECODERS[0], stats_out)                                   # :226
    aux1 = decoder_ref([feature_noise_ref(f, n) for f, n in zip(feats, rand["noise"])],
                       sd, DECODERS[1], stats_out)                                          # :227-228
    aux2 = decoder_ref([dropout_ref(f, m, 0.5) for f, m in zip(feats, rand["aux2_keep"])],
                       sd, DECODERS[2], stats_out)                                          # :229-230
    aux3 = decoder_ref([feature_dropout_ref(f, u) for f, u in zip(feats, rand["u"])],
                       sd, DECODERS[3], stats_out)                                          # :231-232
    return main, aux1, aux2, aux3, feats


def synthetic_state_dict(in_chns: int, class_num: int, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Deterministic weights keyed exactly like ``UNet_UAPS(in_chns, class_num).state_dict()``.

    Values come from a seeded generator in sorted-key order (fan-in scaled normals, BN gamma
    near 1, positive running_var) so fixtures do not depend on nn.Module init order.
    """
    shapes: Dict[str, tuple] = {}

    def block(prefix, cin, cout):
        for idx, (ci, co) in (("0", (cin, cout)), ("4", (cout, cout))):
            shapes[f"{prefix}.conv_conv.{idx}.weight"] = (co, ci, 3, 3)
            shapes[f"{prefix}.conv_conv.{idx}.bias"] = (co,)
        for idx in ("1", "5"):
            for leaf in ("weight", "bias", "running_mean", "running_var"):
                shapes[f"{prefix}.conv_conv.{idx}.{leaf}"] = (cout,)
            shapes[f"{prefix}.conv_conv.{idx}.num_batches_tracked"] = ()

    cin = in_chns
    for lvl in range(5):
        block(ENC_PREFIX[lvl], cin, FT_CHNS[lvl])
        cin = FT_CHNS[lvl]
    for name in DECODERS:
        for i, (c1, c2) in zip((1, 2, 3, 4), ((256, 128), (128, 64), (64, 32), (32, 16))):
            shapes[f"{name}.up{i}.conv1x1.weight"] = (c2, c1, 1, 1)
            shapes[f"{name}.up{i}.conv1x1.bias"] = (c2,)
            block(f"{name}.up{i}.conv", 2 * c2, c2)
        shapes[f"{name}.out_conv.weight"] = (class_num, 16, 3, 3)
        shapes[f"{name}.out_conv.bias"] = (class_num,)

    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key in sorted(shapes):
        shp = shapes[key]
        if key.endswith("num_batches_tracked"):
            sd[key] = torch.zeros((), dtype=torch.long)
        elif key.endswith("running_var"):
            sd[key] = 0.5 + torch.rand(shp, generator=g)
        elif key.endswith("running_mean"):
            sd[key] = 0.1 * torch.randn(shp, generator=g)
        elif len(shp) == 4:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[key] = torch.randn(shp, generator=g) * (2.0 / fan_in) ** 0.5
        elif ".conv_conv.1." in key or ".conv_conv.5." in key:
            sd[key] = (1.0 + 0.1 * torch.randn(shp, generator=g)) if key.endswith("weight") \
                else 0.1 * torch.randn(shp, generator=g)
        else:
            sd[key] = 0.05 * torch.randn(shp, generator=g)
    return sd


def synthetic_rand(feat_shapes: Sequence[tuple], seed: int = 99) -> Dict[str, list]:
    """Injected randomness for one forward; feat_shapes = [(B,C,H,W)] * 5."""
    g = torch.Generator().manual_seed(seed)
    rand = {"enc_keep": [], "noise": [], "aux2_keep": [], "u": []}
    for lvl, shp in enumerate(feat_shapes):
        rand["enc_keep"].append(torch.rand(shp, generator=g) >= ENC_DROPOUT[lvl])
        rand["noise"].append((torch.rand(shp[1:], generator=g) * 2 - 1) * 0.3)
        rand["aux2_keep"].append(torch.rand(shp, generator=g) >= 0.5)
        rand["u"].append(float(0.7 + 0.2 * torch.rand((), generator=g)))
    return rand


def feature_shapes(B: int, H: int, W: int):
    return [(B, c, H >> l, W >> l) for l, c in enumerate(FT_CHNS)]


def unet_layer_trace(x: torch.Tensor, sd: Dict[str, torch.Tensor], rand: Dict[str, list], decoders: Sequence[str] = DECODERS):
    """The forward of ``unet_uaps_ref`` once more, recording every primitive layer as a dict
    {name, kind, inputs, out, ...} with ``retain_grad`` on every tensor, for TEACHER-FORCED per-layer parity tests: each
    hand-written layer of the product is fed the oracle's own input activations (and, backward, the oracle's own upstream
    gradient) and compared with the oracle's output, so rounding does not compound through the 22-layer network.
    kinds: conv (inputs 1 or 2 = the torch.cat halves, :85), bn_act (BatchNorm(train) + LeakyReLU, pre-dropout),
    maxpool, upsample, logits (out_conv)."""
    trace = []

    def keep(t):
        if t.requires_grad:
            t.retain_grad()
        return t

    def rec(**kw):
        trace.append(kw)
        return kw["out"]

    def block(inputs, prefix, p_drop, keep_mask):
        cc = prefix + ".conv_conv"
        xin = inputs[0] if len(inputs) == 1 else torch.cat(list(inputs), dim=1)
        y = keep(F.conv2d(xin, sd[cc + ".0.weight"], sd[cc + ".0.bias"], padding=1))
        rec(name=cc + ".0", kind="conv", inputs=list(inputs), out=y, weight=cc + ".0.weight", bias=cc + ".0.bias", before_bn=True)
        a = keep(F.leaky_relu(_bn_train(y, sd, cc + ".1", None), 0.01))
        rec(name=cc + ".1", kind="bn_act", inputs=[y], out=a, bn=cc + ".1")
        if p_drop > 0.0:
            a = keep(dropout_ref(a, keep_mask, p_drop))
        y2 = keep(F.conv2d(a, sd[cc + ".4.weight"], sd[cc + ".4.bias"], padding=1))
        rec(name=cc + ".4", kind="conv", inputs=[a], out=y2, weight=cc + ".4.weight", bias=cc + ".4.bias", before_bn=True)
        a2 = keep(F.leaky_relu(_bn_train(y2, sd, cc + ".5", None), 0.01))
        rec(name=cc + ".5", kind="bn_act", inputs=[y2], out=a2, bn=cc + ".5")
        return a2

    feats, cur = [], keep(x)
    for lvl in range(5):
        if lvl > 0:
            pooled = keep(F.max_pool2d(cur, 2))
            rec(name=f"pool{lvl}", kind="maxpool", inputs=[cur], out=pooled)
            cur = pooled
        cur = block([cur], ENC_PREFIX[lvl], ENC_DROPOUT[lvl], rand["enc_keep"][lvl])
        feats.append(cur)
    per_dec = {"main_decoder": feats,
               "aux_decoder1": [keep(feature_noise_ref(f, n)) for f, n in zip(feats, rand["noise"])],
               "aux_decoder2": [keep(dropout_ref(f, m, 0.5)) for f, m in zip(feats, rand["aux2_keep"])],
               "aux_decoder3": [keep(feature_dropout_ref(f, u)) for f, u in zip(feats, rand["u"])]}
    outs = []
    for name in decoders:
        fs = per_dec[name]
        cur = fs[4]
        for i, skip in zip((1, 2, 3, 4), (fs[3], fs[2], fs[1], fs[0])):
            up = f"{name}.up{i}"
            y = keep(F.conv2d(cur, sd[up + ".conv1x1.weight"], sd[up + ".conv1x1.bias"]))
            rec(name=up + ".conv1x1", kind="conv", inputs=[cur], out=y, weight=up + ".conv1x1.weight", bias=up + ".conv1x1.bias",
                before_bn=False)
            u2 = keep(F.interpolate(y, scale_factor=2, mode="bilinear", align_corners=True))
            rec(name=up + ".upsample", kind="upsample", inputs=[y], out=u2)
            cur = block([skip, u2], up + ".conv", 0.0, None)
        z = keep(F.conv2d(cur, sd[name + ".out_conv.weight"], sd[name + ".out_conv.bias"], padding=1))
        rec(name=name + ".out_conv", kind="logits", inputs=[cur], out=z, weight=name + ".out_conv.weight",
            bias=name + ".out_conv.bias", before_bn=False)
        outs.append(z)
    return outs, trace
