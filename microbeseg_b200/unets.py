"""Host-side mirror of the reference's network operators, executing on libmbseg (sm_100a CUDA).

Same public surface as /root/reference/src/utils/unets.py: ``build_unet`` (:8-57), ``get_weights``
(:60-78), ``DUNet`` (:380-506) / ``UNet`` (:267-377).  The modules own their parameters in the
reference's layout, so ``state_dict()`` / ``load_state_dict()`` use exactly the published key
grammar (``encoderConv.{i}.conv.{0,2,3,5}.*``, ``pooling.{i}.conv_pool.{0,2}.*``,
``decoder{1,2}Upconv.{i}.{up.0,norm}.*``, ``decoder{1,2}Conv.{i}...``) and a reference ``.pth``
loads unchanged.  ``forward`` does not run torch.nn: it drives the hand-written kernels through
the C ABI (include/mbseg.h): NHWC bf16 activations, fp32 accumulation in TMEM, eval-mode
BatchNorm folded into the conv epilogue, ``torch.cat`` expressed as a second K source, the 1x1
head fused into the last conv.  No fallback: anything the CUDA path does not cover raises.
"""
import ctypes
import os

import torch
import torch.nn as nn

from . import _native as nat

_ACT_CODES = {"relu": 1, "leakyrelu": 2, "elu": 3, "mish": 4}
_IN_CODES = {torch.uint8: 0, torch.uint16: 1, torch.int16: 1, torch.float32: 2}
BN_EPS = 1e-5


class Mish(nn.Module):
    """x * tanh(softplus(x)) (unets.py:81-89); parameter-free placeholder in the module tree."""

    def forward(self, x):
        return x * torch.tanh(nn.functional.softplus(x))


def _act_module(act_fun):
    table = {"relu": lambda: nn.ReLU(inplace=True), "leakyrelu": lambda: nn.LeakyReLU(inplace=True),
             "elu": lambda: nn.ELU(inplace=True), "mish": Mish}
    if act_fun not in table:
        raise Exception('Unsupported activation function: {}'.format(act_fun))
    return table[act_fun]()


def _norm_module(normalization, ch):
    if normalization == 'bn':
        return nn.BatchNorm2d(ch)
    if normalization == 'gn':
        return nn.GroupNorm(num_groups=8, num_channels=ch)
    if normalization == 'in':
        return nn.InstanceNorm2d(num_features=ch)
    raise Exception('Unsupported normalization: {}'.format(normalization))


class ConvBlock(nn.Module):
    """conv3x3 -> act -> norm -> conv3x3 -> act -> norm; parameters under ``conv.{0,2,3,5}``."""

    def __init__(self, ch_in, ch_out, act_fun, normalization):
        super().__init__()
        self.conv = nn.Sequential(
            nn.Conv2d(ch_in, ch_out, kernel_size=3, stride=1, padding=1, bias=True), _act_module(act_fun),
            _norm_module(normalization, ch_out),
            nn.Conv2d(ch_out, ch_out, kernel_size=3, stride=1, padding=1, bias=True), _act_module(act_fun),
            _norm_module(normalization, ch_out))


class ConvPool(nn.Module):
    """conv3x3 stride 2 -> act -> norm; parameters under ``conv_pool.{0,2}``."""

    def __init__(self, ch_in, act_fun, normalization):
        super().__init__()
        self.conv_pool = nn.Sequential(nn.Conv2d(ch_in, ch_in, kernel_size=3, stride=2, padding=1, bias=True),
                                       _act_module(act_fun), _norm_module(normalization, ch_in))


class TranspConvBlock(nn.Module):
    """ConvTranspose2d(2, stride 2) -> norm (no activation); parameters under ``up.0`` / ``norm``."""

    def __init__(self, ch_in, ch_out, normalization):
        super().__init__()
        self.up = nn.Sequential(nn.ConvTranspose2d(ch_in, ch_out, kernel_size=2, stride=2))
        self.norm = _norm_module(normalization, ch_out)


def _channel_plan(filters):
    chans = [int(filters[0])]
    while chans[-1] < filters[1]:
        chans.append(chans[-1] * 2)
    return chans


class _NetBase(nn.Module):
    decoder_names = ()

    def __init__(self, ch_in, ch_out, pool_method, act_fun, normalization, filters):
        super().__init__()
        self.ch_in, self.ch_out, self.filters, self.pool_method = ch_in, ch_out, filters, pool_method
        self.act_fun, self.normalization = act_fun, normalization
        chans = _channel_plan(filters)
        self._chans = chans
        self.encoderConv = nn.ModuleList()
        if pool_method == 'max':
            self.pooling = nn.MaxPool2d(kernel_size=2, stride=2)
        elif pool_method == 'conv':
            self.pooling = nn.ModuleList()
        for i, c in enumerate(chans):
            self.encoderConv.append(ConvBlock(ch_in if i == 0 else chans[i - 1], c, act_fun, normalization))
            if pool_method == 'conv' and i < len(chans) - 1:
                self.pooling.append(ConvPool(c, act_fun, normalization))
        rev = chans[::-1]
        for name in self.decoder_names:
            up, conv = nn.ModuleList(), nn.ModuleList()
            for i in range(len(rev) - 1):
                up.append(TranspConvBlock(rev[i], rev[i + 1], normalization))
                conv.append(ConvBlock(rev[i], rev[i + 1], act_fun, normalization))
            # last 1x1 convolution; the second path of the DU net always has one channel (unets.py:460-461)
            conv.append(nn.Conv2d(rev[-1], 1 if name == "decoder2" else ch_out, kernel_size=1, stride=1, padding=0))
            setattr(self, name + "Upconv", up)
            setattr(self, name + "Conv", conv)
        self._engine = None
        self._engine_key = None

    # -- CUDA engine ---------------------------------------------------------------------------
    def _param_version(self):
        # raw_write_epoch: the training kernels (BatchNorm running statistics, fused optimizer steps) write through
        # raw pointers and do not bump torch's version counters
        return (nat.raw_write_epoch(),) + tuple(
            int(t._version) for t in list(self.parameters()) + list(self.buffers())) + tuple(
            t.data_ptr() for t in list(self.parameters()) + list(self.buffers()))

    def engine(self):
        """Packed-weight execution plan; rebuilt when parameters change (load_state_dict, .to())."""
        key = self._param_version()
        if self._engine is None or self._engine_key != key:
            self._engine = _Engine(self)
            self._engine_key = key
        return self._engine

    # -- fp32 check mode -------------------------------------------------------------------------
    @torch.no_grad()
    def forward_check_fp32(self, x):
        """The eval-mode forward with every convolution as a plain CUDA-core fp32 direct convolution (``mbs_conv_ref_f32``:
        no bf16, no tensor cores) through the SAME layer plan as the product engine.  Diagnostic only (~1 s per 2048^2
        frame): against the fp32 oracle it checks the plan to ~1e-6 at full frame size; the bf16 product path compared
        with it shows the precision policy alone (SURVEY.md 8(c) "fp32 check mode").  BatchNorm networks only.
        ``x``: normalised [N,1,H,W] float CUDA tensor.  Returns the list of head maps [N,ch,H,W] float32."""
        if self.training:
            raise RuntimeError("forward_check_fp32 is the eval-mode forward")
        if not x.is_cuda or self.normalization != 'bn' or self.pool_method != 'conv':
            raise NotImplementedError("fp32 check mode: CUDA tensors, pool_method 'conv', normalization 'bn'")
        L = nat.lib()
        act = _ACT_CODES[self.act_fun]
        n, _, H, W = x.shape
        nl = len(self._chans)
        if H % (1 << (nl - 1)) or W % (1 << (nl - 1)):
            raise RuntimeError(f"model input {H}x{W} is not divisible by {1 << (nl - 1)}")
        dev = x.device

        def conv(mode, src0, src1, weight, bias, bn, a, h, w):
            if mode == 2:
                wt = weight.detach().float().permute(2, 3, 0, 1).reshape(4, weight.shape[0], weight.shape[1]).contiguous()
                cout, ho, wo = weight.shape[1], 2 * h, 2 * w
            else:
                k = weight.shape[2]
                wt = weight.detach().float().permute(2, 3, 1, 0).reshape(k * k, weight.shape[1], weight.shape[0]).contiguous()
                cout = weight.shape[0]
                ho, wo = (h // 2, w // 2) if mode == 1 else (h, w)
            d = nat.ConvRefDesc()
            d.mode, d.N, d.H, d.W = mode, n, h, w
            d.src0, d.C0 = src0.data_ptr(), src0.shape[-1]
            d.src1, d.C1 = (src1.data_ptr(), src1.shape[-1]) if src1 is not None else (None, 0)
            d.weight, d.Cout = wt.data_ptr(), cout
            b = bias.detach().float().contiguous()
            d.bias = b.data_ptr()
            keep = [wt, b]
            if bn is not None:
                scale = (bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
                shift = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
                d.scale, d.shift = scale.data_ptr(), shift.data_ptr()
                keep += [scale, shift]
            else:
                d.scale, d.shift = None, None
            d.act = a
            out = torch.empty((n, ho, wo, cout), dtype=torch.float32, device=dev)
            d.dst = out.data_ptr()
            nat.check(L.mbs_conv_ref_f32(ctypes.byref(d), nat.stream_ptr()), "conv_ref_f32")
            return out

        with torch.cuda.device(dev):
            cur = x.float().permute(0, 2, 3, 1).contiguous()                 # NHWC, C = 1
            skips = []
            for l in range(nl):
                blk = self.encoderConv[l].conv
                cur = conv(0, cur, None, blk[0].weight, blk[0].bias, blk[2], act, H >> l, W >> l)
                cur = conv(0, cur, None, blk[3].weight, blk[3].bias, blk[5], act, H >> l, W >> l)
                skips.append(cur)
                if l < nl - 1:
                    pl = self.pooling[l].conv_pool
                    cur = conv(1, cur, None, pl[0].weight, pl[0].bias, pl[2], act, H >> l, W >> l)
            outs = []
            for name in self.decoder_names:
                ups, convs = getattr(self, name + "Upconv"), getattr(self, name + "Conv")
                y = skips[nl - 1]
                for i in range(nl - 1):
                    l = nl - 2 - i
                    up = conv(2, y, None, ups[i].up[0].weight, ups[i].up[0].bias, ups[i].norm, 0, H >> (l + 1), W >> (l + 1))
                    blk = convs[i].conv
                    y = conv(0, up, skips[l], blk[0].weight, blk[0].bias, blk[2], act, H >> l, W >> l)     # cat([up, skip], 1)
                    y = conv(0, y, None, blk[3].weight, blk[3].bias, blk[5], act, H >> l, W >> l)
                head = convs[nl - 1]
                o = conv(4, y, None, head.weight, head.bias, None, 0, H, W)
                outs.append(o.permute(0, 3, 1, 2).contiguous())
            torch.cuda.current_stream(dev).synchronize()                      # the temporaries above stay alive until here
        return outs

    def _check_supported(self):
        if self.pool_method not in ('conv', 'max') or self.normalization not in ('bn', 'gn', 'in'):
            raise NotImplementedError(
                "the CUDA path covers pool_method 'conv' / 'max' and normalization 'bn' / 'gn' / 'in'; got "
                f"pool_method={self.pool_method!r}, normalization={self.normalization!r}")
        if self.ch_in != 1 or self._chans[0] % 8 != 0:
            raise NotImplementedError("the CUDA path needs ch_in == 1 and filters[0] a multiple of 8")
        if self.normalization != 'bn' and self._chans[0] % 64 != 0:
            raise NotImplementedError("group / instance normalisation needs filters[0] to be a multiple of 64")

    def _forward_maps(self, x):
        if self.training:
            raise RuntimeError("microbeseg_b200: net(x) is the eval-mode forward (call net.eval()); the training step runs through "
                               "microbeseg_b200.training.TrainEngine, not through autograd")
        if not x.is_cuda:
            raise RuntimeError("microbeseg_b200: the network runs on CUDA only (no CPU fallback); got a CPU tensor")
        if x.dim() != 4 or x.shape[1] != 1:
            raise RuntimeError(f"expected input [N,1,H,W], got {tuple(x.shape)}")
        self._check_supported()
        n, _, h, w = x.shape
        return self.engine().run(x.reshape(n, h, w).contiguous().float(), 0, 0, 1.0, 0.0)  # hi < lo: pass through

    def _forward_frame_maps(self, img, pads, lo, hi, lohi_dev):
        """Fused entry used by the frame loop: raw (H,W) uint8/uint16/float32 CUDA frame -> head maps of the
        padded size; normalisation ``2*(x-lo)/(hi-lo)-1`` and top/left padding with ``lo`` happen inside the
        first kernel.  ``lohi_dev`` (float32[2] CUDA tensor from ``frame_minmax``) keeps min/max on the device."""
        self._check_supported()
        if self.training:
            raise RuntimeError("microbeseg_b200: net(x) is the eval-mode forward (call net.eval()); the training step runs through "
                               "microbeseg_b200.training.TrainEngine, not through autograd")
        return self.engine().run(img[None], int(pads[0]), int(pads[1]), float(lo), float(hi), lohi_dev=lohi_dev)


class DUNet(_NetBase):
    """U-net with two decoder paths; returns (x1 border/neighbour map, x2 cell map)  (unets.py:380-506)."""
    decoder_names = ("decoder1", "decoder2")

    def __init__(self, ch_in=1, ch_out=1, pool_method='conv', act_fun='relu', normalization='bn', filters=(64, 1024)):
        super().__init__(ch_in, ch_out, pool_method, act_fun, normalization, filters)

    def forward(self, x):
        outs = self._forward_maps(x)
        return outs[0], outs[1]

    def forward_frame(self, img, pads, lo=0.0, hi=0.0, lohi_dev=None):
        """(border, cell) float32 [1,1,Hp,Wp] for one raw frame (see _forward_frame_maps)."""
        outs = self._forward_frame_maps(img, pads, lo, hi, lohi_dev)
        return outs[0], outs[1]


class UNet(_NetBase):
    """Single-decoder U-net (unets.py:267-377)."""
    decoder_names = ("decoder",)

    def __init__(self, ch_in=1, ch_out=1, pool_method='conv', act_fun='relu', normalization='bn', filters=(64, 1024)):
        super().__init__(ch_in, ch_out, pool_method, act_fun, normalization, filters)

    def forward(self, x):
        return self._forward_maps(x)[0]        # [N, ch_out, H, W] linear outputs (softmax is applied by the caller)

    def forward_frame(self, img, pads, lo=0.0, hi=0.0, lohi_dev=None):
        return self._forward_frame_maps(img, pads, lo, hi, lohi_dev)[0]


def frame_minmax(img, out=None, scratch=None):
    """(min, max) of a raw CUDA frame as a float32[2] CUDA tensor (np.min/np.max of infer_script_local.py:124)."""
    if img.dtype not in _IN_CODES:
        raise RuntimeError(f"unsupported frame dtype {img.dtype}")
    if out is None:
        out = torch.empty(2, dtype=torch.float32, device=img.device)
    if scratch is None:
        scratch = torch.empty(2, dtype=torch.int32, device=img.device)
    img = img.contiguous()
    with torch.cuda.device(img.device):
        nat.check(nat.lib().mbs_frame_minmax(img.data_ptr(), _IN_CODES[img.dtype], img.numel(), out.data_ptr(),
                                             scratch.data_ptr(), nat.stream_ptr()), "frame_minmax")
    return out


def _pad64(c):
    return (int(c) + 63) // 64 * 64


class _Engine:
    """Flat list of kernel launches for one network instance (eval mode)."""

    def __init__(self, net):
        self.L = nat.lib()
        self.net = net
        # the tensor-core kernels work on 64-channel groups: narrower levels (the reference's low-memory fallbacks
        # filters = [32, 512] / [32, 256], train.py:283-288) run zero-padded to 64 channels -- padded output channels
        # get zero weights / bias / shift (act(0) = 0 for every supported activation) and padded input channels zero
        # weight columns, so the real channels are unchanged
        self.chans = [_pad64(c) for c in net._chans]
        self.act = _ACT_CODES[net.act_fun]
        self.device = next(net.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("microbeseg_b200: move the network to a CUDA device first (no CPU fallback)")
        self.bufs = {}
        self.p = {}
        self.norms = {}        # layer name -> (groups, gamma, beta, eps) for the per-sample normalisations ('gn' / 'in')
        self._norm_scratch = None
        with torch.no_grad(), torch.cuda.device(self.device):
            self._pack()

    # -- parameter packing ---------------------------------------------------------------------
    def _affine(self, bn, name=None):
        """eval-mode BatchNorm folds into the conv epilogue; GroupNorm / InstanceNorm2d need the sample's own
        statistics, so the conv gets an identity affine and the layer is normalised in a second pass (_apply_norm)"""
        if isinstance(bn, nn.BatchNorm2d):
            scale = (bn.weight.float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
            shift = (bn.bias.float() - bn.running_mean.float() * scale).contiguous()
            return scale, shift
        if isinstance(bn, nn.GroupNorm):
            c = bn.num_channels
            self.norms[name] = (bn.num_groups, bn.weight.detach().float().contiguous(), bn.bias.detach().float().contiguous(), bn.eps)
        elif isinstance(bn, nn.InstanceNorm2d):
            c = bn.num_features
            if bn.affine or bn.track_running_stats:
                raise NotImplementedError("InstanceNorm2d with affine / running statistics is not what the reference builds")
            self.norms[name] = (c, None, None, bn.eps)
        else:
            raise NotImplementedError(f"unsupported normalisation module {type(bn).__name__}")
        return (torch.ones(c, dtype=torch.float32, device=self.device), torch.zeros(c, dtype=torch.float32, device=self.device))

    def _apply_norm(self, name, t):
        """in-place GroupNorm / InstanceNorm2d of every sample of the NHWC tensor ``t``"""
        groups, gamma, beta, eps = self.norms[name]
        n, h, w, c = t.shape
        if self._norm_scratch is None:
            self._norm_scratch = torch.empty(int(self.L.mbs_bn_scratch_floats(2048)), dtype=torch.float32, device=self.device)
        for b in range(n):
            nat.check(self.L.mbs_sample_group_norm(t[b].data_ptr(), h * w, c, groups, gamma.data_ptr() if gamma is not None else None,
                                                   beta.data_ptr() if beta is not None else None, eps, t[b].data_ptr(),
                                                   self._norm_scratch.data_ptr(), nat.stream_ptr()), "sample_group_norm")

    def _pad_vec(self, v, n, fill):
        v = v.detach().float()
        if v.numel() == n:
            return v.contiguous()
        out = torch.full((n,), float(fill), dtype=torch.float32, device=self.device)
        out[:v.numel()] = v
        return out

    def _pack_conv(self, name, conv, bn, in_parts=None):
        """in_parts: real channel counts of the concatenated sources (each is padded to 64 separately)"""
        w = conv.weight.detach().float()
        cout_r, cin_r = w.shape[0], w.shape[1]
        in_parts = in_parts or [cin_r]
        cout = _pad64(cout_r)
        cin = cin_r if cin_r == 1 else sum(_pad64(c) for c in in_parts)
        if cout != cout_r or cin != cin_r:
            wp = torch.zeros((cout, cin, 3, 3), dtype=torch.float32, device=self.device)
            src = dst = 0
            for c in in_parts:
                wp[:cout_r, dst:dst + c] = w[:, src:src + c]
                src += c
                dst += c if cin_r == 1 else _pad64(c)
            w = wp
        w = w.contiguous()
        if cin == 1:
            packed = w.reshape(cout, 9).contiguous()
        else:
            packed = torch.empty((cout, 9, cin), dtype=torch.bfloat16, device=self.device)
            nat.check(self.L.mbs_pack_conv3x3_weight(w.data_ptr(), cout, cin, packed.data_ptr(), nat.stream_ptr()))
        scale, shift = self._affine(bn, name)
        self.p[name] = (packed, self._pad_vec(conv.bias, cout, 0.0), self._pad_vec(scale, cout, 1.0),
                        self._pad_vec(shift, cout, 0.0), cin, cout)

    def _pack_convT(self, name, block):
        w = block.up[0].weight.detach().float()   # [Cin, Cout, 2, 2]
        cin_r, cout_r = w.shape[0], w.shape[1]
        cin, cout = _pad64(cin_r), _pad64(cout_r)
        if cin != cin_r or cout != cout_r:
            wp = torch.zeros((cin, cout, 2, 2), dtype=torch.float32, device=self.device)
            wp[:cin_r, :cout_r] = w
            w = wp
        w = w.contiguous()
        packed = torch.empty((4 * cout, cin), dtype=torch.bfloat16, device=self.device)
        nat.check(self.L.mbs_pack_convT2x2_weight(w.data_ptr(), cin, cout, packed.data_ptr(), nat.stream_ptr()))
        scale, shift = self._affine(block.norm, name)
        self.p[name] = (packed, self._pad_vec(block.up[0].bias, cout, 0.0), self._pad_vec(scale, cout, 1.0),
                        self._pad_vec(shift, cout, 0.0), cin, cout)

    def _pack(self):
        net = self.net
        for i, blk in enumerate(net.encoderConv):
            self._pack_conv(f"enc{i}a", blk.conv[0], blk.conv[2])
            self._pack_conv(f"enc{i}b", blk.conv[3], blk.conv[5])
        if net.pool_method == 'conv':
            for i, pl in enumerate(net.pooling):
                self._pack_conv(f"pool{i}", pl.conv_pool[0], pl.conv_pool[2])
        for name in net.decoder_names:
            ups, convs = getattr(net, name + "Upconv"), getattr(net, name + "Conv")
            for i, up in enumerate(ups):
                self._pack_convT(f"{name}up{i}", up)
                half = convs[i].conv[0].weight.shape[1] // 2             # cat([up, skip], 1): two equal parts
                self._pack_conv(f"{name}c{i}a", convs[i].conv[0], convs[i].conv[2], in_parts=[half, half])
                self._pack_conv(f"{name}c{i}b", convs[i].conv[3], convs[i].conv[5])
            head = convs[len(ups)]
            if head.weight.shape[0] > 4:
                raise NotImplementedError("the fused 1x1 head supports at most 4 output channels")
            hw = head.weight.detach().float().reshape(head.weight.shape[0], -1)
            if hw.shape[1] != self.chans[0]:
                hp = torch.zeros((hw.shape[0], self.chans[0]), dtype=torch.float32, device=self.device)
                hp[:, :hw.shape[1]] = hw
                hw = hp
            self.p[name + "head"] = (hw.contiguous(), [float(b) for b in head.bias.detach().float().cpu()],
                                     head.bias.detach().float().contiguous())

    # -- buffers -------------------------------------------------------------------------------
    def _buf(self, name, shape, dtype=torch.bfloat16):
        t = self.bufs.get(name)
        if t is None or tuple(t.shape) != tuple(shape):
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self.bufs[name] = t
        return t

    # -- launches ------------------------------------------------------------------------------
    def _conv(self, mode, name, n, h, w, src0, src1, dst, head=None, head_out=None, act=None, first=None):
        packed, bias, scale, shift, cin, cout = self.p[name]
        d = nat.ConvDesc()
        d.mode, d.N, d.H, d.W = mode, n, h, w
        c0 = src0.shape[-1]
        d.src0, d.C0, d.ld0, d.coff0 = src0.data_ptr(), c0, c0, 0
        if src1 is not None:
            c1 = src1.shape[-1]
            d.src1, d.C1, d.ld1, d.coff1 = src1.data_ptr(), c1, c1, 0
        else:
            c1 = 0
            d.src1, d.C1, d.ld1, d.coff1 = None, 0, 0, 0
        assert c0 + c1 == cin, (name, c0, c1, cin)
        d.weight, d.Cout = packed.data_ptr(), cout
        d.bias, d.scale, d.shift = bias.data_ptr(), scale.data_ptr(), shift.data_ptr()
        d.act = self.act if act is None else act
        if dst is not None:
            d.dst, d.ldd, d.coffd = dst.data_ptr(), dst.shape[-1], 0
        else:
            d.dst, d.ldd, d.coffd = None, 0, 0
        if head is not None:
            d.head_w, d.head_n, d.head_out = head[0].data_ptr(), len(head[1]), head_out.data_ptr()
            for k, b in enumerate(head[1]):
                d.head_b[k] = b
        else:
            d.head_w, d.head_n, d.head_out = None, 0, None
        if first is not None:      # the network's first layer is produced inside this convolution's kernel
            img, h0, w0, pad_y, pad_x, lo, hi, lohi_dev, w1, b1, sc1, sh1 = first
            nat.check(self.L.mbs_first_conv_halo64(img.data_ptr(), _IN_CODES[img.dtype], n, h0, w0, pad_y, pad_x, lo, hi,
                                                   lohi_dev.data_ptr() if lohi_dev is not None else None,
                                                   w1.data_ptr(), b1.data_ptr(), sc1.data_ptr(), sh1.data_ptr(),
                                                   self.act, ctypes.byref(d), nat.stream_ptr()), name)
        else:
            nat.check(self.L.mbs_conv_gemm(ctypes.byref(d), nat.stream_ptr()), name)
        self.last_conv_launches += 1
        if name in self.norms and dst is not None:
            self._apply_norm(name, dst)

    def run(self, img, pad_y, pad_x, lo, hi, events=None, lohi_dev=None, keep_features=None):
        """img: [N,H,W] CUDA tensor (uint8 / uint16-as-int16 / float32).  hi < lo -> the values are
        already normalised (reference callers pass the normalised float image)."""
        if img.dtype not in _IN_CODES:
            raise RuntimeError(f"unsupported frame dtype {img.dtype}")
        n, h0, w0 = img.shape
        H, W = h0 + pad_y, w0 + pad_x
        nl = len(self.chans)
        div = 1 << (nl - 1)
        if H % div or W % div:
            # the reference fails with a RuntimeError in torch.cat for such sizes (unets.py:492)
            raise RuntimeError(f"model input {H}x{W} is not divisible by {div}; pad it with zero_pad_model_input")
        ch = self.chans
        img = img.contiguous()
        self.last_conv_launches = 0
        with torch.cuda.device(self.device):
            if events is not None:      # bench instrumentation: [start, after first conv, end]
                events[0].record()
            t1 = [self._buf(f"t1_{l}", (n, H >> l, W >> l, ch[l])) for l in range(nl)]
            t2 = [self._buf(f"t2_{l}", (n, H >> l, W >> l, ch[l])) for l in range(nl)]
            skip = [self._buf(f"skip_{l}", (n, H >> l, W >> l, ch[l])) for l in range(nl)]
            pool = [self._buf(f"pool_{l}", (n, H >> (l + 1), W >> (l + 1), ch[l])) for l in range(nl - 1)]
            # encoder
            packed, bias, scale, shift, _, c0 = self.p["enc0a"]
            # no normalisation layer between the two convolutions of the top block and 64 channels: the first layer is
            # computed inside enc0b's tensor-core kernel (mbs_first_conv_halo64, bit-identical to the two-launch path)
            fused_first = (c0 == 64 and ch[0] == 64 and "enc0a" not in self.norms and "enc0b" not in self.norms
                           and os.environ.get("MBS_FIRST_FUSE", "0") == "1")
            for b in range(0 if fused_first else n):
                nat.check(self.L.mbs_first_conv(img[b].data_ptr(), _IN_CODES[img.dtype], h0, w0, pad_y, pad_x, lo, hi,
                                                lohi_dev.data_ptr() if lohi_dev is not None else None,
                                                packed.data_ptr(), bias.data_ptr(), scale.data_ptr(),
                                                shift.data_ptr(), c0, self.act, t1[0][b].data_ptr(), c0, 0,
                                                nat.stream_ptr()), "first_conv")
            if "enc0a" in self.norms:
                self._apply_norm("enc0a", t1[0])
            if events is not None:
                events[1].record()
            for l in range(nl):
                if l > 0:
                    self._conv(0, f"enc{l}a", n, H >> l, W >> l, pool[l - 1], None, t1[l])
                if l == 0 and fused_first:
                    self._conv(0, "enc0b", n, H, W, t1[0], None, skip[0],
                               first=(img, h0, w0, pad_y, pad_x, lo, hi, lohi_dev, packed, bias, scale, shift))
                else:
                    self._conv(0, f"enc{l}b", n, H >> l, W >> l, t1[l], None, skip[l])
                if l < nl - 1:
                    if self.net.pool_method == 'max':          # nn.MaxPool2d(2, 2), unets.py:363-364
                        nat.check(self.L.mbs_maxpool2x2(skip[l].data_ptr(), n, H >> l, W >> l, ch[l], pool[l].data_ptr(),
                                                        nat.stream_ptr()), "maxpool2x2")
                    else:
                        self._conv(1, f"pool{l}", n, H >> l, W >> l, skip[l], None, pool[l])
            outs = []
            for name in self.net.decoder_names:
                x = skip[nl - 1]
                out = torch.empty((n, len(self.p[name + "head"][1]), H, W), dtype=torch.float32, device=self.device)
                for i in range(nl - 1):
                    l = nl - 2 - i
                    self._conv(2, f"{name}up{i}", n, H >> (l + 1), W >> (l + 1), x, None, t1[l], act=0)
                    self._conv(0, f"{name}c{i}a", n, H >> l, W >> l, t1[l], skip[l], t2[l])
                    if l > 0:
                        self._conv(0, f"{name}c{i}b", n, H >> l, W >> l, t2[l], None, t1[l])
                        x = t1[l]
                    else:
                        feat = None
                        if keep_features is not None:      # calibration only: also write the last 64-channel map
                            feat = torch.empty((n, H, W, ch[0]), dtype=torch.bfloat16, device=self.device)
                            keep_features[name] = feat
                        lname = f"{name}c{i}b"
                        if lname in self.norms:
                            # the normalisation sits between the last conv and the 1x1 head: no fused head
                            if feat is None:
                                feat = self._buf("feat_last", (n, H, W, ch[0]))
                            self._conv(0, lname, n, H, W, t2[0], None, feat)
                            hw, hb, hb_dev = self.p[name + "head"]
                            for b in range(n):
                                for k in range(len(hb)):
                                    nat.check(self.L.mbs_head_fwd(feat[b].data_ptr(), H * W, ch[0], hw[k].data_ptr(),
                                                                  hb_dev[k:k + 1].data_ptr(), out[b, k].data_ptr(),
                                                                  nat.stream_ptr()), "head_fwd")
                        else:
                            self._conv(0, lname, n, H, W, t2[0], None, feat, head=self.p[name + "head"][:2], head_out=out)
                outs.append(out)
            if events is not None:
                events[2].record()
        return outs


def build_unet(unet_type, act_fun, pool_method, normalization, device, num_gpus, ch_in=1, ch_out=1,
               filters=(64, 1024)):
    """Build U-net architecture (same signature as unets.py:8-57).

    ``num_gpus > 1`` meant nn.DataParallel in the reference (training only, unets.py:51-52); here
    multi-GPU inference shards frames over one process per GPU (microbeseg_b200.inference), so the
    module itself always lives on one device."""
    if unet_type == 'DU':
        model = DUNet(ch_in=ch_in, ch_out=ch_out, pool_method=pool_method, filters=filters, act_fun=act_fun,
                      normalization=normalization)
    elif unet_type == 'U':
        model = UNet(ch_in=ch_in, ch_out=ch_out, pool_method=pool_method, filters=filters, act_fun=act_fun,
                     normalization=normalization)
    else:
        raise Exception('Architecture "{}" is not known'.format(unet_type))
    return model.to(device)


def get_weights(net, weights, device, num_gpus):
    """Load a reference ``state_dict`` ``.pth`` into the model (same signature as unets.py:60-78)."""
    state = torch.load(weights, map_location=device)
    target = net.module if hasattr(net, "module") else net
    target.load_state_dict(state)
    return net
