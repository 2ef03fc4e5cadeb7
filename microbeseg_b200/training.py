"""Training step of the distance U-Net on the CUDA path (no autograd, no cuDNN).

Mirrors the numerical part of the reference's inner loop (src/training/train.py:460-493):
``zero_grad -> net(img) -> SmoothL1(border) + SmoothL1(cell) -> backward -> optimizer.step`` for the published
configuration (DU net, conv pooling, BatchNorm, ReLU, ``smooth_l1``).  Forward and backward convolutions run on
the tcgen05 kernels of libmbseg (forward / data-gradient: ``mbs_conv_gemm``; weight-gradient: ``mbs_conv_wgrad``);
BatchNorm (batch statistics), ReLU', the 1x1 heads and the loss are the kernels of ``train_kernels.cu``.
Activations and activation gradients are bf16 (fp32 accumulation), parameters / parameter gradients fp32
("bf16 autocast with fp32 master weights").  The optimizer is the reference's own ``torch.optim.Adam`` and the
gradient exchange is one NCCL all-reduce over a flat fp32 buffer (replaces nn.DataParallel, unets.py:51-52);
BatchNorm statistics stay per rank, as with DataParallel.
"""
import ctypes

import torch

from . import _native as nat

BN_EPS = 1e-5
BN_MOMENTUM = 0.1
ACT_NONE, ACT_RELU, ACT_MISH = 0, 1, 4
_LOSS_KINDS = {'smooth_l1': 0, 'l1': 1, 'l2': 2}      # get_loss(loss_function, 'distance'), losses.py:24-35
_BOUNDARY_LOSSES = {'ce': 0, 'ce_dice': 1}           # get_loss(loss_function, 'boundary'), losses.py:16-21 (value = with_dice)


class GradBuckets:
    """Flat fp32 gradient buffer laid out in the order the backward pass PRODUCES the gradients, cut into buckets of
    about ``bucket_bytes`` (~25 MB).  ``views[p]`` is the slice a parameter's gradient lives in (the
    kernels write there directly, so there is no flatten / unflatten copy); bucket k can be all-reduced as soon as the
    parameter ``closing[k]`` has its gradient, while the backward pass carries on with the earlier layers."""

    def __init__(self, params_in_order, device, bucket_bytes=25 << 20):
        self.params = list(params_in_order)
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.views, self.closing, self.ranges = {}, {}, []
        off = start = acc = 0
        for p in self.params:
            n = p.numel()
            self.views[p] = self.flat[off:off + n].view(p.shape)
            off += n
            acc += 4 * n
            if acc >= bucket_bytes:
                self.closing[p] = len(self.ranges)
                self.ranges.append((start, off))
                start, acc = off, 0
        if start < total:
            self.closing[self.params[-1]] = len(self.ranges)
            self.ranges.append((start, total))
        self._works = []
        self._comm = None

    def allreduce(self, k, world_size):
        """asynchronous mean over the ranks of bucket k, on a side stream that waits for the work enqueued so far"""
        import torch.distributed as dist
        s, e = self.ranges[k]
        buf = self.flat[s:e]
        if buf.is_cuda:
            cur = torch.cuda.current_stream(buf.device)
            if self._comm is None:
                self._comm = torch.cuda.Stream(buf.device)
            ev = torch.cuda.Event()
            ev.record(cur)
            with torch.cuda.stream(self._comm):
                self._comm.wait_event(ev)
                if dist.get_backend() == "nccl":
                    w = dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True)
                    self._works.append((w, None))
                else:
                    w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)
                    self._works.append((w, (buf, world_size)))
        else:                                   # CPU tensors (gloo): tests of the host logic
            w = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=True)
            self._works.append((w, (buf, world_size)))

    def finish(self):
        """the current stream waits for every outstanding bucket (no host synchronisation on CUDA)"""
        for w, post in self._works:
            w.wait()
            if post is not None:
                post[0].div_(post[1])
        self._works = []


def _pad64(c):
    return (int(c) + 63) // 64 * 64


class _Padded:
    """Zero-padded fp32 twin of a parameter / buffer whose channel dimensions are not multiples of 64 (the reference's
    low-memory fallback nets, filters = [32, 512] / [32, 256], train.py:283-288).  The tensor-core kernels work on
    64-channel groups; padded output channels have zero weights / bias / beta, so their activations, their gradients
    and every gradient that touches a padded input channel are EXACT zeros and the real region is sliced back out."""
    __slots__ = ("src", "buf", "index", "writeback")

    def __init__(self, src, dims, fill, writeback=False):
        shape = list(src.shape)
        per_dim = {}
        for dim, parts in dims:
            idx, off = [], 0
            for c in parts:
                idx.append(torch.arange(c, device=src.device) + off)
                off += _pad64(c)
            per_dim[dim] = torch.cat(idx)
            shape[dim] = off
        index, n_idx, k = [], len(per_dim), 0
        for d in range(max(per_dim) + 1):
            if d in per_dim:
                view = [1] * n_idx                    # broadcastable index tensors: the result keeps the dimension order
                view[k] = -1
                index.append(per_dim[d].view(view))
                k += 1
            else:
                index.append(slice(None))
        self.src, self.index, self.writeback = src, tuple(index), writeback
        self.buf = torch.full(shape, float(fill), dtype=torch.float32, device=src.device)


class _Layer:
    __slots__ = ("name", "kind", "conv", "bn", "act", "srcs", "a", "y", "mean", "invstd", "geom", "x_f32")


class TrainEngine:
    """Forward + backward of one network on one GPU; fills ``param.grad`` (fp32) for every parameter.
    Distance method: DUNet with loss 'smooth_l1' | 'l1' | 'l2' (two regression heads); boundary method: UNet with three
    output classes and loss 'ce' | 'ce_dice' (losses.py:16-21, 71-96)."""

    def __init__(self, net, use_graph=True, loss='smooth_l1'):
        from .unets import DUNet, UNet
        self.boundary = loss in _BOUNDARY_LOSSES
        if self.boundary:
            if not isinstance(net, UNet) or net.ch_out != 3:
                raise NotImplementedError("loss 'ce' / 'ce_dice' trains the boundary method: a 'U' network with ch_out=3")
        elif not isinstance(net, DUNet):
            raise NotImplementedError("the distance criteria ('smooth_l1', 'l1', 'l2') train the DU (distance) network")
        net._check_supported()
        if net.pool_method != 'conv' or net.normalization != 'bn':
            raise NotImplementedError("the CUDA training step needs pool_method 'conv' and normalization 'bn' "
                                      "(what the reference's TrainWorker builds, train.py:184-188)")
        if net.act_fun not in ("relu", "mish"):
            raise NotImplementedError("training on the CUDA path supports act_fun='relu' (Adam recipe) and 'mish' "
                                      "(Ranger recipe), the two the reference trains with (train.py:174)")
        # relu: the conv epilogue applies it and the stored tensor is post-activation (its sign is the derivative mask);
        # mish: the conv writes the PRE-activation z, the BatchNorm passes apply mish on load and the backward pass
        # evaluates mish'(z) -- still one activation-sized tensor per layer
        self.act = ACT_RELU if net.act_fun == "relu" else ACT_MISH
        self.conv_act = ACT_RELU if net.act_fun == "relu" else ACT_NONE
        if loss not in _LOSS_KINDS and loss not in _BOUNDARY_LOSSES:
            raise Exception('Loss unknown')                 # get_loss, losses.py:33-34
        self.loss_kind = _BOUNDARY_LOSSES[loss] + 16 if self.boundary else _LOSS_KINDS[loss]
        self.net = net
        self.L = nat.lib()
        self.dev = next(net.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("microbeseg_b200.training needs a CUDA device (no CPU fallback)")
        self.chans = net._chans
        self._const = {}
        self._scratch = torch.empty(int(self.L.mbs_bn_scratch_floats(2048)), dtype=torch.float32, device=self.dev)
        self.use_graph = use_graph
        self._graphs = {}
        # gradient buckets (built after the first step has shown the production order) and the data-parallel exchange
        self._order, self._buckets = [], None
        self.world_size = 1
        self._capture = None
        self._packs = None
        # narrow levels (filters[0] = 32: the reference's out-of-memory fallbacks) run zero-padded to 64 channels
        self._pad = {}
        if any(c % 64 for c in self.chans):
            self._build_padding()

    # ---- channel padding of the narrow nets ---------------------------------------------------------
    def _build_padding(self):
        net, ch = self.net, self.chans
        nl = len(ch)

        def reg(t, dims, fill=0.0, writeback=False):
            if any(_pad64(c) != c for _, parts in dims for c in parts):
                self._pad[t] = _Padded(t, dims, fill, writeback)

        def reg_bn(bn, c):
            reg(bn.weight, [(0, [c])], 1.0)
            reg(bn.bias, [(0, [c])])
            reg(bn.running_mean, [(0, [c])], 0.0, True)
            reg(bn.running_var, [(0, [c])], 1.0, True)

        def reg_conv(conv, cout, in_parts):
            reg(conv.weight, [(0, [cout])] + ([(1, in_parts)] if in_parts else []))
            reg(conv.bias, [(0, [cout])])

        for l in range(nl):
            blk = net.encoderConv[l]
            reg_conv(blk.conv[0], ch[l], [ch[l - 1]] if l else None)
            reg_bn(blk.conv[2], ch[l])
            reg_conv(blk.conv[3], ch[l], [ch[l]])
            reg_bn(blk.conv[5], ch[l])
            if l < nl - 1:
                reg_conv(net.pooling[l].conv_pool[0], ch[l], [ch[l]])
                reg_bn(net.pooling[l].conv_pool[2], ch[l])
        for name in net.decoder_names:
            ups, convs = getattr(net, name + "Upconv"), getattr(net, name + "Conv")
            for i in range(nl - 1):
                l = nl - 2 - i
                up = ups[i].up[0]                                   # ConvTranspose2d weight [Cin][Cout][2][2]
                reg(up.weight, [(0, [ch[l + 1]]), (1, [ch[l]])])
                reg(up.bias, [(0, [ch[l]])])
                reg_bn(ups[i].norm, ch[l])
                reg_conv(convs[i].conv[0], ch[l], [ch[l], ch[l]])      # torch.cat([up, skip]): each source padded on its own
                reg_bn(convs[i].conv[2], ch[l])
                reg_conv(convs[i].conv[3], ch[l], [ch[l]])
                reg_bn(convs[i].conv[5], ch[l])
            reg(convs[nl - 1].weight, [(1, [ch[0]])])

    def _v(self, t):
        """the tensor the kernels see: the zero-padded twin of a narrow parameter / buffer, else the tensor itself"""
        pd = self._pad.get(t)
        return t.detach() if pd is None else pd.buf

    def _refresh_padded(self):
        for pd in self._pad.values():
            pd.buf[pd.index] = pd.src.detach()

    def _writeback_padded(self):
        for pd in self._pad.values():
            if pd.writeback:
                pd.src.copy_(pd.buf[pd.index])

    # ---- small helpers ------------------------------------------------------------------------
    def _ones(self, c):
        k = ("1", c)
        if k not in self._const:
            self._const[k] = torch.ones(c, dtype=torch.float32, device=self.dev)
        return self._const[k]

    def _zeros(self, c):
        k = ("0", c)
        if k not in self._const:
            self._const[k] = torch.zeros(c, dtype=torch.float32, device=self.dev)
        return self._const[k]

    def _sp(self):
        return nat.stream_ptr()

    # ---- gradient placement / data-parallel exchange ---------------------------------------------
    def _grad_buf(self, param):
        """where the kernel that produces ``param``'s gradient writes it: its slice of the flat bucket buffer"""
        pd = self._pad.get(param)
        if pd is not None:                              # narrow layer: the kernel fills a padded buffer, _set_grad slices it
            return torch.empty(pd.buf.shape, dtype=torch.float32, device=self.dev)
        if self._buckets is not None:
            return self._buckets.views[param]
        return torch.empty(param.shape, dtype=torch.float32, device=self.dev)

    def _set_grad(self, param, tensor):
        pd = self._pad.get(param)
        if pd is not None:
            tensor = tensor.reshape(pd.buf.shape)[pd.index]
        if self._buckets is None:                       # first (eager) step: remember the production order
            param.grad = tensor.reshape(param.shape).contiguous()
            self._order.append(param)
            return
        view = self._buckets.views[param]
        if tensor.data_ptr() != view.data_ptr():
            view.copy_(tensor.reshape(view.shape))
        param.grad = view
        k = self._buckets.closing.get(param)
        if k is not None:
            self._boundary(k)

    def _boundary(self, k):
        """bucket k is complete.  While capturing: close the current CUDA graph and open the next one (the all-reduce
        is launched between the graph replays, never captured); otherwise launch the bucket's all-reduce now."""
        cap = self._capture
        if cap is not None:
            cap["cur"].capture_end()
            cap["graphs"].append((cap["cur"], k))
            if k + 1 < len(self._buckets.ranges):
                cap["cur"] = torch.cuda.CUDAGraph()
                cap["cur"].capture_begin(pool=cap["pool"])
            else:
                cap["cur"] = None                     # the last gradient of the step: nothing is launched after it
        elif self.world_size > 1:
            self._buckets.allreduce(k, self.world_size)

    def finish_allreduce(self):
        if self._buckets is not None:
            self._buckets.finish()

    def _pack_all(self):
        """GEMM-packed bf16 copies of every conv / transposed-conv weight (forward and data-gradient forms) in ONE
        launch per step; the packed buffers are persistent, the job table is rebuilt when a weight's storage moves."""
        net = self.net
        convs = []
        for m in net.modules():
            if isinstance(m, torch.nn.Conv2d) and m.kernel_size == (3, 3) and m.in_channels > 1:
                convs.append((m, 0))
            elif isinstance(m, torch.nn.ConvTranspose2d):
                convs.append((m, 1))
        key = tuple(self._v(m.weight).data_ptr() for m, _ in convs)
        if self._packs is None or self._packs["key"] != key:
            jobs = (nat.PackJob * len(convs))()
            bufs, tile0 = {}, 0
            for i, (m, kind) in enumerate(convs):
                w = self._v(m.weight)
                if w.dtype != torch.float32 or not w.is_contiguous():
                    raise RuntimeError("training needs contiguous fp32 parameters")
                cout, cin = (w.shape[0], w.shape[1]) if kind == 0 else (w.shape[1], w.shape[0])
                if cout % 32 or cin % 32:
                    raise RuntimeError("training needs conv channel counts that are multiples of 32")
                taps = 9 if kind == 0 else 4
                fwd = torch.empty((cout, 9, cin) if kind == 0 else (4 * cout, cin), dtype=torch.bfloat16, device=self.dev)
                dg = torch.empty((cin, taps, cout), dtype=torch.bfloat16, device=self.dev)
                bufs[m] = (fwd, dg)
                j = jobs[i]
                j.w, j.fwd, j.dgrad = w.data_ptr(), fwd.data_ptr(), dg.data_ptr()
                j.cout, j.cin, j.kind, j.tile0 = cout, cin, kind, tile0
                tile0 += (cout // 32) * (cin // 32)
            table = torch.frombuffer(bytearray(bytes(jobs)), dtype=torch.uint8).to(self.dev)
            self._packs = {"key": key, "bufs": bufs, "table": table, "n": len(convs), "tiles": tile0}
        pk = self._packs
        nat.check(self.L.mbs_pack_train_weights(pk["table"].data_ptr(), pk["n"], pk["tiles"], self._sp()), "pack_train_weights")

    def _conv(self, mode, n, h, w, srcs, packed, cout, bias, act, dst):
        d = nat.ConvDesc()
        d.mode, d.N, d.H, d.W = mode, n, h, w
        s0 = srcs[0]
        d.src0, d.C0, d.ld0, d.coff0 = s0.data_ptr(), s0.shape[-1], s0.shape[-1], 0
        if len(srcs) > 1:
            s1 = srcs[1]
            d.src1, d.C1, d.ld1, d.coff1 = s1.data_ptr(), s1.shape[-1], s1.shape[-1], 0
        else:
            d.src1, d.C1, d.ld1, d.coff1 = None, 0, 0, 0
        d.weight, d.Cout = packed.data_ptr(), cout
        d.bias, d.scale, d.shift = bias.data_ptr(), self._ones(cout).data_ptr(), self._zeros(cout).data_ptr()
        d.act = act
        d.dst, d.ldd, d.coffd = dst.data_ptr(), dst.shape[-1], 0
        d.head_w, d.head_n, d.head_out = None, 0, None
        nat.check(self.L.mbs_conv_gemm(ctypes.byref(d), self._sp()), "conv_gemm(train)")

    def _bn_fwd(self, lay):
        a, bn = lay.a, lay.bn
        c = a.shape[-1]
        m = a.numel() // c
        lay.y = torch.empty_like(a)
        lay.mean = torch.empty(c, dtype=torch.float32, device=self.dev)
        lay.invstd = torch.empty(c, dtype=torch.float32, device=self.dev)
        # batch statistics, y = BN(a), and the running statistics (torch defaults: momentum 0.1, unbiased variance)
        nat.check(self.L.mbs_bn_train_fwd(a.data_ptr(), m, c, self._v(bn.weight).data_ptr(), self._v(bn.bias).data_ptr(), BN_EPS,
                                          lay.y.data_ptr(), self._scratch.data_ptr(), lay.mean.data_ptr(),
                                          lay.invstd.data_ptr(), BN_MOMENTUM, self._v(bn.running_mean).data_ptr(),
                                          self._v(bn.running_var).data_ptr(), bn.num_batches_tracked.data_ptr(),
                                          ACT_MISH if lay.act == ACT_MISH else ACT_NONE, self._sp()),
                  "bn_train_fwd")

    def _wgrad(self, kind, n, ho, wo, a, cm, b, cn):
        """weight gradient on the tensor cores straight from the NHWC bf16 activations: ``a`` = dz (conv) / d(up)
        (transposed conv), ``b`` = the layer input.  Deterministic: every K split stores its fp32 tile into a scratch
        buffer [splits][cm][taps][cn] (returned with the split count); ``_wgrad_reduce`` sums the splits in order."""
        d = nat.WgradDesc()
        d.kind, d.N, d.Ho, d.Wo = kind, n, ho, wo
        d.a, d.Cm, d.lda, d.coffa = a.data_ptr(), cm, a.shape[-1], 0
        d.b, d.Cn, d.ldb, d.coffb = b.data_ptr(), cn, b.shape[-1], 0
        d.out, d.out_ld, d.out_coff, d.partial = None, cn, 0, 1
        splits = int(self.L.mbs_conv_wgrad_splits(ctypes.byref(d)))
        if splits <= 0:
            nat.check(1, "conv_wgrad_splits")
        part = torch.empty((splits, cm, 4 if kind == 2 else 9, cn), dtype=torch.float32, device=self.dev)
        d.out = part.data_ptr()
        nat.check(self.L.mbs_conv_wgrad(ctypes.byref(d), self._sp()), "conv_wgrad")
        return part, splits

    def _wgrad_reduce(self, parts, cm, layout, out):
        (p0, s0), (p1, s1) = parts[0], (parts[1] if len(parts) > 1 else (None, 0))
        nat.check(self.L.mbs_wgrad_reduce(p0.data_ptr(), s0, p0.shape[-1], p1.data_ptr() if p1 is not None else None, s1,
                                          p1.shape[-1] if p1 is not None else 0, cm, layout, out.data_ptr(), self._sp()), "wgrad_reduce")

    # ---- forward ------------------------------------------------------------------------------
    def _fwd_conv(self, name, conv, bn, srcs, stride=1):
        lay = _Layer()
        lay.name, lay.kind, lay.conv, lay.bn, lay.act, lay.srcs = name, ("s2" if stride == 2 else "s1"), conv, bn, self.act, srcs
        n, h, w = srcs[0].shape[:3]
        cout = self._v(conv.weight).shape[0]
        ho, wo = (h // 2, w // 2) if stride == 2 else (h, w)
        lay.a = torch.empty((n, ho, wo, cout), dtype=torch.bfloat16, device=self.dev)
        self._conv(1 if stride == 2 else 0, n, h, w, srcs, self._packs["bufs"][conv][0], cout,
                   self._v(conv.bias).float(), self.conv_act, lay.a)
        self._bn_fwd(lay)
        self.tape.append(lay)
        return lay

    def _fwd_first(self, conv, bn, x):
        lay = _Layer()
        lay.name, lay.kind, lay.conv, lay.bn, lay.act, lay.srcs, lay.x_f32 = "enc0a", "first", conv, bn, self.act, [], x
        n, h, w = x.shape
        c = self._v(conv.weight).shape[0]
        lay.a = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=self.dev)
        wt = self._v(conv.weight).float().reshape(c, 9).contiguous()
        b = self._v(conv.bias).float().contiguous()
        for i in range(n):
            nat.check(self.L.mbs_first_conv(x[i].data_ptr(), 2, h, w, 0, 0, 1.0, 0.0, None, wt.data_ptr(), b.data_ptr(),
                                            self._ones(c).data_ptr(), self._zeros(c).data_ptr(), c, self.conv_act,
                                            lay.a[i].data_ptr(), c, 0, self._sp()), "first_conv(train)")
        self._bn_fwd(lay)
        self.tape.append(lay)
        return lay

    def _fwd_up(self, name, block, x):
        lay = _Layer()
        conv = block.up[0]
        lay.name, lay.kind, lay.conv, lay.bn, lay.act, lay.srcs = name, "up", conv, block.norm, ACT_NONE, [x]
        n, h, w, cin = x.shape
        cout = self._v(conv.weight).shape[1]
        packed = self._packs["bufs"][conv][0]
        lay.a = torch.empty((n, 2 * h, 2 * w, cout), dtype=torch.bfloat16, device=self.dev)
        self._conv(2, n, h, w, [x], packed, cout, self._v(conv.bias).float(), ACT_NONE, lay.a)
        self._bn_fwd(lay)
        self.tape.append(lay)
        return lay

    def forward_backward(self, img, border_label, cell_label=None, world_size=1):
        """img / labels: [N,1,H,W] float32 CUDA tensors (img normalised to [-1,1] as the reference's ToTensor does).
        Boundary method: ``border_label`` is the [N,H,W] class image (0 / 1 / 2) and ``cell_label`` stays None.
        Returns the loss (0-d tensor) and leaves the gradients in ``param.grad``.

        The first call of a given batch shape runs eagerly (it warms up every kernel configuration and records the
        order in which the backward pass produces the gradients); gradients then live in one flat buffer cut into
        ~25 MB buckets.  With ``use_graph`` the second call captures the step -- ~600 kernel launches plus the torch
        glue -- into one CUDA graph PER BUCKET, and later calls replay them; with ``world_size`` > 1 the all-reduce of
        bucket k is launched (on a side stream) right after graph k, so it overlaps the rest of the backward pass.
        Call ``finish_allreduce()`` (``train_step`` does) before the optimizer reads the gradients."""
        nat.note_raw_write()        # BatchNorm running statistics are updated through raw pointers
        self.world_size = int(world_size)
        first = self._buckets is None
        if first or not self.use_graph:
            loss = self._forward_backward(img, border_label, cell_label)
            if first:
                self._buckets = GradBuckets(self._order, self.dev)
                if self.world_size > 1:                 # this one step: gradients are not in the flat buffer yet
                    allreduce_gradients(self.net, self.world_size)
            return loss
        # the captured graphs bake in parameter / buffer storage and the loss kind: re-capture when any of them moves
        # (net.to(), load_state_dict(assign=True), p.data rebinding)
        key = (tuple(img.shape), img.dtype, self.loss_kind) + tuple(
            t.data_ptr() for t in list(self.net.parameters()) + list(self.net.buffers()))
        st = self._graphs.get(key)
        if st is None:
            st = {"in": [None if t is None else torch.empty_like(t, dtype=torch.float32) for t in (img, border_label, cell_label)]}
            for d, s_ in zip(st["in"], (img, border_label, cell_label)):
                if d is not None:
                    d.copy_(s_)
            self._pack_all()                    # (re)builds the packed-weight buffers and their job table outside the capture
            torch.cuda.synchronize(self.dev)
            self.L.mbs_launch_count(1)
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            cap = {"graphs": [], "pool": torch.cuda.graph_pool_handle(), "cur": torch.cuda.CUDAGraph()}
            with torch.cuda.stream(side):
                cap["cur"].capture_begin(pool=cap["pool"])
                self._capture = cap
                try:
                    st["loss"] = self._forward_backward(*st["in"])
                finally:
                    self._capture = None
                    if cap["cur"] is not None:
                        cap["cur"].capture_end()
                if cap["cur"] is not None:
                    cap["graphs"].append((cap["cur"], None))
            torch.cuda.current_stream(self.dev).wait_stream(side)
            st["graphs"] = cap["graphs"]
            st["launches"] = int(self.L.mbs_launch_count(0))     # kernels of this library inside one replay of all graphs
            # the capture does not execute the step; the gradient views are re-attached after every replay
            # (optimizer.zero_grad(set_to_none=True) detaches them)
            st["grads"] = [(p, self._buckets.views[p]) for p in self.net.parameters()]
            self._graphs = {key: st}                             # one live capture (its private pool holds every activation)
        for d, s_ in zip(st["in"], (img, border_label, cell_label)):
            if d is not None:
                d.copy_(s_)
        for g, k in st["graphs"]:
            g.replay()
            if k is not None and self.world_size > 1:
                self._buckets.allreduce(k, self.world_size)
        for p, gr in st["grads"]:
            p.grad = gr
        self.launches_last_step = st["launches"]
        return st["loss"].clone()       # the static loss tensor is overwritten by the next replay

    def _forward_backward(self, img, border_label, cell_label):
        net = self.net
        if not net.training:
            raise RuntimeError("call net.train() before a training step")
        n, _, H, W = img.shape
        nl = len(self.chans)
        if H % (1 << (nl - 1)) or W % (1 << (nl - 1)):
            raise RuntimeError(f"training crops must be divisible by {1 << (nl - 1)}")
        self.tape = []
        with torch.cuda.device(self.dev), torch.no_grad():
            x = img.reshape(n, H, W).contiguous().float()
            self._refresh_padded()
            self._pack_all()
            # ---------------- forward ----------------
            enc_a, enc_b, pools = [], [], []
            cur = None
            for l in range(nl):
                blk = net.encoderConv[l]
                la = self._fwd_first(blk.conv[0], blk.conv[2], x) if l == 0 else \
                    self._fwd_conv(f"enc{l}a", blk.conv[0], blk.conv[2], [cur])
                lb = self._fwd_conv(f"enc{l}b", blk.conv[3], blk.conv[5], [la.y])
                enc_a.append(la)
                enc_b.append(lb)
                if l < nl - 1:
                    pl = net.pooling[l]
                    lp = self._fwd_conv(f"pool{l}", pl.conv_pool[0], pl.conv_pool[2], [lb.y], stride=2)
                    pools.append(lp)
                    cur = lp.y
            dec = {}
            preds, heads, last = [], [], []
            targets = [border_label, cell_label]
            for name in net.decoder_names:
                ups, convs = getattr(net, name + "Upconv"), getattr(net, name + "Conv")
                xcur = enc_b[nl - 1].y
                chain = []
                for i in range(nl - 1):
                    l = nl - 2 - i
                    lu = self._fwd_up(f"{name}up{i}", ups[i], xcur)
                    la = self._fwd_conv(f"{name}c{i}a", convs[i].conv[0], convs[i].conv[2], [lu.y, enc_b[l].y])
                    lb = self._fwd_conv(f"{name}c{i}b", convs[i].conv[3], convs[i].conv[5], [la.y])
                    chain.append((lu, la, lb, l))
                    xcur = lb.y
                dec[name] = chain
                head = convs[nl - 1]
                n_out = head.weight.shape[0]
                hw = self._v(head.weight).float().reshape(n_out, -1).contiguous()
                pred = torch.empty((n_out, n, H, W), dtype=torch.float32, device=self.dev)      # planar per output channel
                for k in range(n_out):
                    nat.check(self.L.mbs_head_fwd(xcur.data_ptr(), n * H * W, hw.shape[1], hw[k].data_ptr(), head.bias[k:k + 1].data_ptr(),
                                                  pred[k].data_ptr(), self._sp()), "head_fwd")
                preds.append(pred)
                heads.append((head, hw))
                last.append(xcur)
            # ---------------- loss ----------------
            loss = torch.zeros(1, dtype=torch.float32, device=self.dev)
            gpred = []
            if self.boundary:
                # one 'U' decoder with three class logits; label image with classes 0 / 1 / 2 (train.py:468-470, 483-484)
                lab8 = border_label.reshape(n, H, W).to(torch.uint8).contiguous()
                g = torch.empty_like(preds[0])
                sums = torch.empty(7, dtype=torch.float64, device=self.dev)
                nat.check(self.L.mbs_ce_dice_loss(preds[0].data_ptr(), lab8.data_ptr(), n * H * W, self.loss_kind - 16, loss.data_ptr(),
                                                  g.data_ptr(), sums.data_ptr(), self._sp()), "ce_dice_loss")
                gpred.append(g)
            else:
                for pred, tgt in zip(preds, targets):
                    g = torch.empty_like(pred)
                    t = tgt.reshape(n, H, W).contiguous().float()
                    nat.check(self.L.mbs_regression_loss(pred.data_ptr(), t.data_ptr(), n * H * W, self.loss_kind, loss.data_ptr(),
                                                         g.data_ptr(), self._sp()), "regression_loss")
                    gpred.append(g)
            # ---------------- backward ----------------
            skip_grads = [[] for _ in range(nl)]          # contributions to d(skip_l)
            bott_grads = []
            for di, name in enumerate(net.decoder_names):
                head, hw = heads[di]
                y_last = last[di]
                n_out, c0 = hw.shape
                dys, dws = [], []
                for k in range(n_out):
                    dyk = torch.empty_like(y_last)
                    dwdb = torch.empty(c0 + 1, dtype=torch.float32, device=self.dev)
                    nat.check(self.L.mbs_head_bwd(gpred[di][k].data_ptr(), y_last.data_ptr(), n * H * W, c0, hw[k].data_ptr(),
                                                  dyk.data_ptr(), dwdb.data_ptr(), self._scratch.data_ptr(), self._sp()), "head_bwd")
                    dys.append(dyk)
                    dws.append(dwdb)
                dy = self._sum(dys)
                if n_out == 1:
                    self._set_grad(head.weight, dws[0][:c0])
                    self._set_grad(head.bias, dws[0][c0:])
                else:
                    stacked = torch.stack(dws)
                    self._set_grad(head.weight, stacked[:, :c0].contiguous())
                    self._set_grad(head.bias, stacked[:, c0].contiguous())
                for (lu, la, lb, l) in reversed(dec[name]):
                    dy = self._bwd_conv(lb, dy)[0]
                    dup, dskip = self._bwd_conv(la, dy)
                    skip_grads[l].append(dskip)
                    dy = self._bwd_up(lu, dup)
                bott_grads.append(dy)
            dy = self._sum(bott_grads)
            for l in range(nl - 1, -1, -1):
                dy = self._bwd_conv(enc_b[l], dy)[0]
                if l == 0:
                    self._bwd_first(enc_a[0], dy)
                    break
                dy = self._bwd_conv(enc_a[l], dy)[0]          # gradient w.r.t. pool_{l-1}.y
                dsk = self._bwd_conv(pools[l - 1], dy)[0]      # ... w.r.t. skip_{l-1}
                dy = self._sum(skip_grads[l - 1] + [dsk])
            self._writeback_padded()
        self.tape = []
        return loss[0]

    # ---- backward pieces ----------------------------------------------------------------------
    def _sum(self, ts):
        if len(ts) == 1:
            return ts[0]
        out = torch.empty_like(ts[0])
        c = ts[2] if len(ts) > 2 else None
        nat.check(self.L.mbs_add3_bf16(ts[0].data_ptr(), ts[1].data_ptr(), c.data_ptr() if c is not None else None,
                                       ts[0].numel(), out.data_ptr(), self._sp()), "add3")
        assert len(ts) <= 3
        return out

    def _bn_bwd(self, lay, dy):
        a, bn = lay.a, lay.bn
        c = a.shape[-1]
        m = a.numel() // c
        dy = dy.contiguous()
        dz = torch.empty_like(a)
        dgb = torch.empty(2 * c, dtype=torch.float32, device=self.dev)
        dbias = torch.empty(c, dtype=torch.float32, device=self.dev)
        nat.check(self.L.mbs_bn_train_bwd(dy.data_ptr(), a.data_ptr(), m, c, lay.mean.data_ptr(), lay.invstd.data_ptr(),
                                          self._v(bn.weight).data_ptr(), lay.act, dz.data_ptr(), dgb.data_ptr(), dbias.data_ptr(),
                                          self._scratch.data_ptr(), self._sp()), "bn_train_bwd")
        self._set_grad(bn.weight, dgb[:c])
        self._set_grad(bn.bias, dgb[c:])
        self._set_grad(lay.conv.bias, dbias)
        return dz

    def _bwd_conv(self, lay, dy):
        """conv3x3 (stride 1 or 2) + ReLU + BN.  Returns the input gradients, one per source."""
        conv = lay.conv
        dz = self._bn_bwd(lay, dy)
        n, ho, wo, cout = dz.shape
        cins = [s.shape[-1] for s in lay.srcs]
        cin = sum(cins)
        stride2 = lay.kind == "s2"
        # weight gradient: per-split tiles dW[k][co][tap][ci] per source -> ordered sum in the reference layout [Cout,Cin,3,3]
        assert len(lay.srcs) <= 2
        parts = [self._wgrad(1 if stride2 else 0, n, ho, wo, dz, cout, s, cs) for s, cs in zip(lay.srcs, cins)]
        dw = self._grad_buf(conv.weight)
        self._wgrad_reduce(parts, cout, 0, dw)
        self._set_grad(conv.weight, dw)
        # data gradient: full-resolution stride-1 conv with the flipped, transposed filter [Cin][9][Cout]
        packed = self._packs["bufs"][conv][1]
        src = dz
        h, w = ho, wo
        if stride2:
            src = torch.empty((n, 2 * ho, 2 * wo, cout), dtype=torch.bfloat16, device=self.dev)
            nat.check(self.L.mbs_zero_insert_up2(dz.data_ptr(), n, ho, wo, cout, src.data_ptr(), self._sp()), "zero_insert")
            h, w = 2 * ho, 2 * wo
        # one conv per source of a concatenated input: the rows [off, off+cs) of the packed filter are that source's
        # gradient, written contiguously (no strided split copies)
        dxs, off = [], 0
        for cs in cins:
            dx = torch.empty((n, h, w, cs), dtype=torch.bfloat16, device=self.dev)
            self._conv(0, n, h, w, [src], packed[off:off + cs], cs, self._zeros(cs), ACT_NONE, dx)
            dxs.append(dx)
            off += cs
        return dxs

    def _bwd_first(self, lay, dy):
        conv = lay.conv
        dz = self._bn_bwd(lay, dy)
        n, h, w, c = dz.shape
        dw = torch.empty((c, 9), dtype=torch.float32, device=self.dev)
        nat.check(self.L.mbs_first_conv_wgrad(lay.x_f32.data_ptr(), dz.data_ptr(), n, h, w, c, dw.data_ptr(),
                                              self._scratch.data_ptr(), self._sp()),
                  "first_conv_wgrad")
        self._set_grad(conv.weight, dw)

    def _bwd_up(self, lay, dy):
        """ConvTranspose2d(2,2) + BN (no activation): returns the gradient w.r.t. its input."""
        conv = lay.conv
        dup = self._bn_bwd(lay, dy)                        # [N, 2H, 2W, Cout]
        x = lay.srcs[0]
        n, h, w, cin = x.shape
        cout = dup.shape[-1]
        dw = self._grad_buf(conv.weight)                   # [cin, cout, 2, 2]
        self._wgrad_reduce([self._wgrad(2, n, h, w, dup, cout, x, cin)], cout, 1, dw)
        self._set_grad(conv.weight, dw)
        # data gradient = 2x2 stride-2 convolution of d(up) with W[ci][co][q]
        packed = self._packs["bufs"][conv][1]
        dx = torch.empty((n, h, w, cin), dtype=torch.bfloat16, device=self.dev)
        self._conv(3, n, 2 * h, 2 * w, [dup], packed, cin, self._zeros(cin), ACT_NONE, dx)
        return dx


def flat_grads(net):
    return [p for p in net.parameters() if p.grad is not None]


def allreduce_gradients(net, world_size):
    """Data-parallel gradient exchange: ONE all-reduce over a flat fp32 buffer, then the mean (NCCL over NVLink)."""
    import torch.distributed as dist
    params = flat_grads(net)
    grads = [p.grad for p in params]
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(world_size)
    views, off = [], 0
    for g in grads:
        k = g.numel()
        views.append(flat[off:off + k].view_as(g))
        off += k
    torch._foreach_copy_(grads, views)          # one multi-tensor launch instead of 270 small copies
    return flat.numel()


DDP_DESCRIPTION = ("gradients in one flat buffer in backward order, ~25 MB buckets, NCCL all-reduce (AVG) of bucket k launched "
                   "after CUDA graph k on a side stream: overlaps the rest of the backward pass")


def broadcast_module_state(net, src=0):
    """Replicas start from rank ``src``'s parameters and BatchNorm buffers, like nn.DataParallel's replicate
    (unets.py:51-52).  During training every rank keeps its own batch statistics (DataParallel semantics); call this
    again before evaluating / saving so that rank 0's running statistics are the ones that survive, as in the reference."""
    import torch.distributed as dist
    for t in list(net.parameters()) + list(net.buffers()):
        dist.broadcast(t.data, src)
    nat.note_raw_write()


def train_step(engine, optimizer, img, border_label, cell_label=None, world_size=1):
    """One optimisation step as in train.py:473-493 (zero_grad, forward, loss, backward, step)."""
    optimizer.zero_grad(set_to_none=True)
    loss = engine.forward_backward(img, border_label, cell_label, world_size)
    engine.finish_allreduce()
    optimizer.step()
    return loss
