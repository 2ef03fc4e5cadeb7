"""Adam / AMSGrad with the whole step fused into ONE CUDA launch.

Drop-in for ``torch.optim.Adam`` as the reference constructs it (src/training/train.py:380-385: lr 8e-4, betas
(0.9, 0.999), eps 1e-8, weight_decay 0, amsgrad=True): same constructor arguments, same ``state`` entries (``step``,
``exp_avg``, ``exp_avg_sq``, ``max_exp_avg_sq``), same arithmetic as torch's single-tensor path.  torch's multi-tensor
implementation makes ~8 passes over the 46 M parameters of the published net (1.5 ms on a B200); ``mbs_adam_step``
makes one (5 fp32 streams in, 4 out).  fp32 parameters on a CUDA device only (no CPU fallback).
"""
import torch
from torch.optim.optimizer import Optimizer

from . import _native as nat

_CHUNK = 1024


class Adam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if not 0.0 <= eps:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        if not 0.0 <= weight_decay:
            raise ValueError(f"Invalid weight_decay value: {weight_decay}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, amsgrad=amsgrad))
        self._tables = {}
        self.launches_last_step = 0

    def _table(self, plist, amsgrad):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg'].data_ptr()) for p in plist) + (amsgrad,)
        hit = self._tables.get(key)
        if hit is not None:
            return hit
        arr = (nat.RangerTensor * len(plist))()
        row = 0
        for i, p in enumerate(plist):
            st = self.state[p]
            e = arr[i]
            e.p, e.g = p.data_ptr(), p.grad.data_ptr()
            e.exp_avg, e.exp_avg_sq = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
            e.slow = st['max_exp_avg_sq'].data_ptr() if amsgrad else st['exp_avg_sq'].data_ptr()
            e.numel = p.numel()
            e.rows, e.row_len, e.gc = (p.numel() + _CHUNK - 1) // _CHUNK, _CHUNK, 0
            e.row_start = row
            row += e.rows
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(plist[0].device)
        self._tables = {key: (raw, row)}          # keep only the current table
        return self._tables[key]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        L = nat.lib()
        self.launches_last_step = 0
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            buckets = {}
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Adam does not support sparse gradients, please consider SparseAdam instead')
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("microbeseg_b200.adam.Adam needs fp32 CUDA parameters (no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("Adam: parameters and gradients must be contiguous")
                state = self.state[p]
                if len(state) == 0:
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p)
                    state['exp_avg_sq'] = torch.zeros_like(p)
                    if group['amsgrad']:
                        state['max_exp_avg_sq'] = torch.zeros_like(p)
                state['step'] += 1
                buckets.setdefault(int(state['step']), []).append(p)
            for step, plist in buckets.items():
                bias_correction1 = 1 - beta1 ** step
                bias_correction2 = 1 - beta2 ** step
                step_size = group['lr'] / bias_correction1
                table, rows = self._table(plist, bool(group['amsgrad']))
                with torch.cuda.device(plist[0].device):
                    nat.check(L.mbs_adam_step(table.data_ptr(), len(plist), rows, beta1, beta2, group['eps'], group['weight_decay'],
                                              step_size, bias_correction2 ** 0.5, 1 if group['amsgrad'] else 0, nat.stream_ptr()),
                              "adam_step")
                self.launches_last_step += 1
        nat.note_raw_write()        # parameters changed behind torch's version counters: cached eval engines are stale
        return loss
