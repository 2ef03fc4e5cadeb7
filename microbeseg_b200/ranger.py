"""Ranger optimizer (RAdam + gradient centralisation + lookahead) with the whole step fused into ONE CUDA launch.

Drop-in for the reference's default optimizer, /root/reference/src/training/ranger2020.py:43-210 (same constructor
arguments and defaults, same ``state`` layout -- ``step``, ``exp_avg``, ``exp_avg_sq``, ``slow_buffer`` -- so optimizer
state dicts interchange).  The reference walks the parameters in Python and issues ~12 small ATen kernels per tensor
(270 tensors for the published net); here the scalar schedule (N_sma, step_size, :165-180) is evaluated on the host
exactly as the reference does, and ``mbs_ranger_step`` updates every tensor in one launch.
fp32 parameters on a CUDA device only (no CPU fallback).
"""
import ctypes
import math

import torch
from torch.optim.optimizer import Optimizer

from . import _native as nat

_CHUNK = 1024          # row length used for 1-D tensors (no centralisation there)


class Ranger(Optimizer):
    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999), eps=1e-5, weight_decay=0,
                 use_gc=True, gc_conv_only=False, gc_loc=True):
        # parameter checks, ranger2020.py:53-60
        if not 0.0 <= alpha <= 1.0:
            raise ValueError(f'Invalid slow update rate: {alpha}')
        if not 1 <= k:
            raise ValueError(f'Invalid lookahead steps: {k}')
        if not lr > 0:
            raise ValueError(f'Invalid Learning Rate: {lr}')
        if not eps > 0:
            raise ValueError(f'Invalid eps: {eps}')
        if not gc_loc:
            raise NotImplementedError("gc_loc=False (centralising the update instead of the gradient) is not built; "
                                      "the reference's callers use the default gc_loc=True")
        defaults = dict(lr=lr, alpha=alpha, k=k, step_counter=0, betas=betas, N_sma_threshhold=N_sma_threshhold, eps=eps,
                        weight_decay=weight_decay)
        super().__init__(params, defaults)
        self.N_sma_threshhold = N_sma_threshhold
        self.alpha = alpha
        self.k = k
        self.radam_buffer = [[None, None, None] for _ in range(10)]
        self.gc_loc = gc_loc
        self.use_gc = use_gc
        self.gc_conv_only = gc_conv_only
        self._tables = {}          # tuple of data pointers -> (device table, host array kept alive, total rows)
        self.launches_last_step = 0

    def _schedule(self, step, beta1, beta2):
        """ranger2020.py:163-180 (memoised per step % 10 like the reference's radam_buffer)."""
        buffered = self.radam_buffer[int(step % 10)]
        if step == buffered[0]:
            return buffered[1], buffered[2]
        buffered[0] = step
        beta2_t = beta2 ** step
        N_sma_max = 2 / (1 - beta2) - 1
        N_sma = N_sma_max - 2 * step * beta2_t / (1 - beta2_t)
        buffered[1] = N_sma
        if N_sma > self.N_sma_threshhold:
            step_size = math.sqrt((1 - beta2_t) * (N_sma - 4) / (N_sma_max - 4) * (N_sma - 2) / N_sma * N_sma_max /
                                  (N_sma_max - 2)) / (1 - beta1 ** step)
        else:
            step_size = 1.0 / (1 - beta1 ** step)
        buffered[2] = step_size
        return N_sma, step_size

    def _gc(self, p):
        if not self.use_gc:
            return False
        return p.dim() > 3 if self.gc_conv_only else p.dim() > 1

    def _table(self, plist):
        key = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg'].data_ptr(), self.state[p]['exp_avg_sq'].data_ptr(),
                     self.state[p]['slow_buffer'].data_ptr()) for p in plist)
        hit = self._tables.get(key)
        if hit is not None:
            return hit
        arr = (nat.RangerTensor * len(plist))()
        row = 0
        for i, p in enumerate(plist):
            st = self.state[p]
            e = arr[i]
            e.p, e.g = p.data_ptr(), p.grad.data_ptr()
            e.exp_avg, e.exp_avg_sq, e.slow = st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr(), st['slow_buffer'].data_ptr()
            e.numel = p.numel()
            if self._gc(p):
                e.rows, e.row_len, e.gc = p.shape[0], p.numel() // p.shape[0], 1
            else:
                e.rows, e.row_len, e.gc = (p.numel() + _CHUNK - 1) // _CHUNK, _CHUNK, 0
            e.row_start = row
            row += e.rows
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).to(plist[0].device)
        self._tables = {key: (raw, row)}          # keep only the current table
        return self._tables[key]

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        L = nat.lib()
        self.launches_last_step = 0
        for group in self.param_groups:
            beta1, beta2 = group['betas']
            buckets = {}                                   # parameters that share the step count share one launch
            for p in group['params']:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError('Ranger optimizer does not support sparse gradients')
                if not p.is_cuda or p.dtype != torch.float32 or p.grad.dtype != torch.float32:
                    raise RuntimeError("microbeseg_b200.ranger.Ranger needs fp32 CUDA parameters (no CPU fallback)")
                if not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError("Ranger: parameters and gradients must be contiguous")
                state = self.state[p]
                if len(state) == 0:                        # ranger2020.py:118-128
                    state['step'] = 0
                    state['exp_avg'] = torch.zeros_like(p)
                    state['exp_avg_sq'] = torch.zeros_like(p)
                    state['slow_buffer'] = torch.empty_like(p)
                    state['slow_buffer'].copy_(p)
                state['step'] += 1
                buckets.setdefault(state['step'], []).append(p)
            for step, plist in buckets.items():
                N_sma, step_size = self._schedule(step, beta1, beta2)
                table, rows = self._table(plist)
                with torch.cuda.device(plist[0].device):
                    nat.check(L.mbs_ranger_step(table.data_ptr(), len(plist), rows, beta1, beta2, group['eps'],
                                                group['weight_decay'], step_size * group['lr'],
                                                1 if N_sma > self.N_sma_threshhold else 0,
                                                1 if step % group['k'] == 0 else 0, self.alpha, nat.stream_ptr()),
                              "ranger_step")
                self.launches_last_step += 1
        nat.note_raw_write()        # parameters changed behind torch's version counters: cached eval engines are stale
        return loss
