"""ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product package).

Plain fp32 PyTorch restatement of the reference network forward pass
(/root/reference/src/utils/unets.py: ConvBlock :92-173, ConvPool :176-226, TranspConvBlock
:229-264, UNet.forward :349-377, DUNet.forward :463-506), written functionally over a reference
layout state dict (key grammar of SURVEY.md 8(a) row A4).

PINNED: tests/golden/net_*.npz hold outputs of the REAL reference module (imported from
/root/reference by tests/golden/make_golden.py) on seeded weights; tests/test_oracle_net.py
checks this restatement against them, so both this oracle and the CUDA path are anchored to
the reference itself.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _act(x, act):
    if act == "relu":
        return F.relu(x)
    if act == "leakyrelu":
        return F.leaky_relu(x, 0.01)
    if act == "elu":
        return F.elu(x)
    if act == "mish":
        return x * torch.tanh(F.softplus(x))       # unets.py:81-89
    raise ValueError(act)


def _bn(x, sd, prefix):
    """the block's normalisation in eval mode (unets.py:127-132): BatchNorm2d if the state dict holds running statistics,
    GroupNorm(8, C) if only weight / bias, InstanceNorm2d (no parameters, no buffers) otherwise"""
    if prefix + ".running_mean" in sd:
        return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"],
                            sd[prefix + ".bias"], training=False, eps=1e-5)
    if prefix + ".weight" in sd:
        return F.group_norm(x, 8, sd[prefix + ".weight"], sd[prefix + ".bias"], eps=1e-5)
    return F.instance_norm(x, eps=1e-5)


# Storage-policy emulation ("check mode").  The CUDA path stores every layer output and every tensor-core weight in
# bf16 and accumulates in fp32.  With _POLICY["bf16"] the oracle rounds exactly those tensors to bf16 (values stay
# fp32 tensors): the weights of every conv with more than one input channel, and every layer output except the last
# 64-channel map, which the fused 1x1 head consumes unrounded.  What remains between this oracle and the CUDA path is
# fp32 summation order (and the bf16 roundings that flip because of it), so a kernel bug shows up far above the
# residual, while the plain fp32 oracle measures the precision policy itself.
_POLICY = {"bf16": False, "keep_next": False}


def _r(x):
    return x.to(torch.bfloat16).float() if _POLICY["bf16"] else x


def _w(w):
    return w.to(torch.bfloat16).float() if (_POLICY["bf16"] and w.shape[1] > 1) else w


def _out(x):
    """a layer output as the CUDA path stores it"""
    if _POLICY["keep_next"]:
        _POLICY["keep_next"] = False
        return x
    return _r(x)


def _conv_block(x, sd, prefix, act, last=False):
    # conv -> act -> norm -> conv -> act -> norm (unets.py:112-160; note: norm AFTER the activation)
    x = _out(_bn(_act(F.conv2d(x, _w(sd[prefix + ".conv.0.weight"]), sd[prefix + ".conv.0.bias"], padding=1), act), sd,
                 prefix + ".conv.2"))
    _POLICY["keep_next"] = last
    x = _out(_bn(_act(F.conv2d(x, _w(sd[prefix + ".conv.3.weight"]), sd[prefix + ".conv.3.bias"], padding=1), act), sd,
                 prefix + ".conv.5"))
    return x


def _conv_pool(x, sd, prefix, act):
    x = F.conv2d(x, _w(sd[prefix + ".conv_pool.0.weight"]), sd[prefix + ".conv_pool.0.bias"], stride=2, padding=1)
    return _out(_bn(_act(x, act), sd, prefix + ".conv_pool.2"))


def _upconv(x, sd, prefix):
    x = F.conv_transpose2d(x, _w(sd[prefix + ".up.0.weight"]), sd[prefix + ".up.0.bias"], stride=2)
    return _out(_bn(x, sd, prefix + ".norm"))      # no activation (unets.py:261-262)


def n_levels(sd):
    return len({k.split(".")[1] for k in sd if k.startswith("encoderConv.")})


def encoder(sd, x, act):
    skips = []
    nl = n_levels(sd)
    for i in range(nl - 1):
        x = _conv_block(x, sd, f"encoderConv.{i}", act)
        skips.append(x)
        if f"pooling.{i}.conv_pool.0.weight" in sd:
            x = _conv_pool(x, sd, f"pooling.{i}", act)
        else:                                      # pool_method 'max': nn.MaxPool2d(2, 2) has no parameters (unets.py:306-307)
            x = F.max_pool2d(x, kernel_size=2, stride=2)
    x = _conv_block(x, sd, f"encoderConv.{nl - 1}", act)
    return x, skips[::-1]


def decoder(sd, x, skips, act, name):
    for i, s in enumerate(skips):
        x = _upconv(x, sd, f"{name}Upconv.{i}")
        x = torch.cat([x, s], 1)                   # up first, then skip (unets.py:492)
        x = _conv_block(x, sd, f"{name}Conv.{i}", act, last=(i == len(skips) - 1))
    k = len(skips)
    return F.conv2d(x, sd[f"{name}Conv.{k}.weight"], sd[f"{name}Conv.{k}.bias"])


@torch.no_grad()
def dunet_forward(sd, x, act="relu", policy="fp32"):
    """DUNet.forward (unets.py:463-506): returns (x1 = border/neighbour map, x2 = cell map).
    ``policy='bf16'``: emulate the CUDA path's bf16 storage (see _POLICY); default = the reference's plain fp32."""
    sd = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
    _POLICY["bf16"], _POLICY["keep_next"] = (policy == "bf16"), False
    try:
        b, skips = encoder(sd, x.float(), act)
        return decoder(sd, b, skips, act, "decoder1"), decoder(sd, b, skips, act, "decoder2")
    finally:
        _POLICY["bf16"], _POLICY["keep_next"] = False, False


@torch.no_grad()
def unet_forward(sd, x, act="relu"):
    """UNet.forward (unets.py:349-377)."""
    sd = {k: v.float() for k, v in sd.items() if v.dtype.is_floating_point}
    b, skips = encoder(sd, x.float(), act)
    return decoder(sd, b, skips, act, "decoder")


def seeded_state_dict(template, seed):
    """Deterministic, torch-version independent weights for a reference-layout state dict.

    He-scaled conv weights, non-trivial BatchNorm statistics (so that the eval-mode affine is
    exercised); every tensor has its own numpy stream keyed by (seed, crc32(key)), so the values do
    not depend on the key order of the template."""
    import zlib
    out = {}
    for k, v in template.items():
        rng = np.random.default_rng([int(seed), zlib.crc32(k.encode())])
        shape = tuple(v.shape)
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(100, dtype=torch.int64)
        elif k.endswith("running_var"):
            out[k] = torch.from_numpy(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif k.endswith("running_mean"):
            out[k] = torch.from_numpy(rng.normal(0, 0.1, shape).astype(np.float32))
        elif len(shape) == 4:
            if ".up." in k:      # ConvTranspose2d weight [Cin, Cout, 2, 2]: each output sees Cin inputs
                fan_in = shape[0]
            else:
                fan_in = shape[1] * shape[2] * shape[3]
            out[k] = torch.from_numpy(rng.normal(0, np.sqrt(2.0 / fan_in), shape).astype(np.float32))
        elif k.endswith(".weight"):   # BatchNorm gamma
            out[k] = torch.from_numpy(rng.uniform(0.7, 1.3, shape).astype(np.float32))
        else:                         # biases / BatchNorm beta
            out[k] = torch.from_numpy(rng.normal(0, 0.05, shape).astype(np.float32))
    return out


def reference_layout_template(unet_type="DU", filters=(64, 1024), ch_in=1, ch_out=1, pool_method="conv", normalization="bn"):
    """Shapes of a reference state dict without importing the reference."""
    t = {}

    def conv(prefix, cin, cout, k):
        t[prefix + ".weight"] = torch.empty(cout, cin, k, k)
        t[prefix + ".bias"] = torch.empty(cout)

    def bn(prefix, c):
        if normalization == "in":                  # nn.InstanceNorm2d default: no parameters, no buffers
            return
        names = ("weight", "bias", "running_mean", "running_var") if normalization == "bn" else ("weight", "bias")
        for n in names:
            t[f"{prefix}.{n}"] = torch.empty(c)
        if normalization == "bn":
            t[prefix + ".num_batches_tracked"] = torch.empty((), dtype=torch.int64)

    def block(prefix, cin, cout):
        conv(prefix + ".conv.0", cin, cout, 3)
        bn(prefix + ".conv.2", cout)
        conv(prefix + ".conv.3", cout, cout, 3)
        bn(prefix + ".conv.5", cout)

    chans = [filters[0]]
    while chans[-1] < filters[1]:
        chans.append(chans[-1] * 2)
    for i, c in enumerate(chans):
        block(f"encoderConv.{i}", ch_in if i == 0 else chans[i - 1], c)
    for i, c in enumerate(chans[:-1]):
        if pool_method == "conv":                  # nn.MaxPool2d has no parameters
            conv(f"pooling.{i}.conv_pool.0", c, c, 3)
            bn(f"pooling.{i}.conv_pool.2", c)
    names = ["decoder1", "decoder2"] if unet_type == "DU" else ["decoder"]
    for name in names:
        rc = chans[::-1]
        for i in range(len(rc) - 1):
            t[f"{name}Upconv.{i}.up.0.weight"] = torch.empty(rc[i], rc[i + 1], 2, 2)
            t[f"{name}Upconv.{i}.up.0.bias"] = torch.empty(rc[i + 1])
            bn(f"{name}Upconv.{i}.norm", rc[i + 1])
        for i in range(len(rc) - 1):
            block(f"{name}Conv.{i}", rc[i], rc[i + 1])
        co = 1 if name == "decoder2" else ch_out
        conv(f"{name}Conv.{len(rc) - 1}", rc[-1], co, 1)
    return t


# ------------------------------------------------------------------------------------------
# training-mode reference (plain PyTorch fp32 with autograd) for the training-step parity tests:
# the same graph as DUNet.forward in train() mode (BatchNorm with batch statistics) followed by
# SmoothL1Loss(border) + SmoothL1Loss(cell)  (src/training/train.py:479-482, losses.py:30-32).
# ------------------------------------------------------------------------------------------
def _bn_train(x, sd, prefix):
    return F.batch_norm(x, None, None, sd[prefix + ".weight"], sd[prefix + ".bias"], training=True, eps=1e-5)


def _conv_block_train(x, sd, prefix, act="relu"):
    x = _bn_train(_act(F.conv2d(x, sd[prefix + ".conv.0.weight"], sd[prefix + ".conv.0.bias"], padding=1), act), sd, prefix + ".conv.2")
    return _bn_train(_act(F.conv2d(x, sd[prefix + ".conv.3.weight"], sd[prefix + ".conv.3.bias"], padding=1), act), sd, prefix + ".conv.5")


def dunet_train_loss(params, x, border_label, cell_label, act="relu", loss="smooth_l1"):
    """params: dict name -> tensor (requires_grad where wanted).  Returns the scalar loss."""
    nl = n_levels(params)
    skips = []
    for i in range(nl - 1):
        x = _conv_block_train(x, params, f"encoderConv.{i}", act)
        skips.append(x)
        x = F.conv2d(x, params[f"pooling.{i}.conv_pool.0.weight"], params[f"pooling.{i}.conv_pool.0.bias"], stride=2, padding=1)
        x = _bn_train(_act(x, act), params, f"pooling.{i}.conv_pool.2")
    b = _conv_block_train(x, params, f"encoderConv.{nl - 1}", act)
    skips = skips[::-1]
    outs = []
    for name in ("decoder1", "decoder2"):
        y = b
        for i, s in enumerate(skips):
            y = F.conv_transpose2d(y, params[f"{name}Upconv.{i}.up.0.weight"], params[f"{name}Upconv.{i}.up.0.bias"], stride=2)
            y = _bn_train(y, params, f"{name}Upconv.{i}.norm")
            y = _conv_block_train(torch.cat([y, s], 1), params, f"{name}Conv.{i}", act)
        k = len(skips)
        outs.append(F.conv2d(y, params[f"{name}Conv.{k}.weight"], params[f"{name}Conv.{k}.bias"]))
    crit = {'smooth_l1': torch.nn.SmoothL1Loss, 'l1': torch.nn.L1Loss, 'l2': torch.nn.MSELoss}[loss]()      # losses.py:24-35
    return crit(outs[0], border_label) + crit(outs[1], cell_label)


def unet_train_loss(params, x, label, act="relu", loss="ce_dice"):
    """boundary method: 'U' net (one decoder, 3 class logits) in train() mode + ce / ce_dice (train.py:483-484)"""
    from . import losses as ol
    nl = n_levels(params)
    skips = []
    for i in range(nl - 1):
        x = _conv_block_train(x, params, f"encoderConv.{i}", act)
        skips.append(x)
        x = F.conv2d(x, params[f"pooling.{i}.conv_pool.0.weight"], params[f"pooling.{i}.conv_pool.0.bias"], stride=2, padding=1)
        x = _bn_train(_act(x, act), params, f"pooling.{i}.conv_pool.2")
    y = _conv_block_train(x, params, f"encoderConv.{nl - 1}", act)
    for i, sk in enumerate(skips[::-1]):
        y = F.conv_transpose2d(y, params[f"decoderUpconv.{i}.up.0.weight"], params[f"decoderUpconv.{i}.up.0.bias"], stride=2)
        y = _bn_train(y, params, f"decoderUpconv.{i}.norm")
        y = _conv_block_train(torch.cat([y, sk], 1), params, f"decoderConv.{i}", act)
    k = len(skips)
    logits = F.conv2d(y, params[f"decoderConv.{k}.weight"], params[f"decoderConv.{k}.bias"])
    return ol.boundary_loss(logits, label, loss)
