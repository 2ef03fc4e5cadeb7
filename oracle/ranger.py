"""ORACLE (test infrastructure only -- never imported by the product path).

NumPy float32 restatement of the reference's Ranger optimizer step,
/root/reference/src/training/ranger2020.py:101-210 (RAdam + gradient centralisation :30-40 + lookahead :204-210).
PINNED: tests/golden/ranger_*.npz are produced by the reference's own ranger2020.py on CPU
(tests/golden/make_golden.py ranger); tests/test_oracle_ranger.py checks this restatement against them.
"""
import math

import numpy as np


class RangerOracle:
    def __init__(self, params, lr=1e-3, alpha=0.5, k=6, N_sma_threshhold=5, betas=(.95, 0.999), eps=1e-5, weight_decay=0,
                 use_gc=True, gc_conv_only=False):
        self.p = [np.array(a, np.float32) for a in params]
        self.m = [np.zeros_like(a) for a in self.p]
        self.v = [np.zeros_like(a) for a in self.p]
        self.slow = [a.copy() for a in self.p]
        self.step_no = 0
        self.lr, self.alpha, self.k, self.thr, self.betas, self.eps, self.wd = lr, alpha, k, N_sma_threshhold, betas, eps, weight_decay
        self.use_gc, self.gc_conv_only = use_gc, gc_conv_only

    def schedule(self, step):          # :165-180
        beta1, beta2 = self.betas
        beta2_t = beta2 ** step
        n_max = 2 / (1 - beta2) - 1
        n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
        if n_sma > self.thr:
            ss = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) / (1 - beta1 ** step)
        else:
            ss = 1.0 / (1 - beta1 ** step)
        return n_sma, ss

    def step(self, grads):
        f = np.float32
        beta1, beta2 = self.betas
        self.step_no += 1
        n_sma, ss = self.schedule(self.step_no)
        out_g = []
        for i, g in enumerate(grads):
            g = np.array(g, np.float32)
            if self.use_gc and (g.ndim > 3 if self.gc_conv_only else g.ndim > 1):          # :30-40
                g = g + (-g.mean(axis=tuple(range(1, g.ndim)), keepdims=True, dtype=np.float32))
            out_g.append(g)
            self.v[i] = self.v[i] * f(beta2) + f(1 - beta2) * g * g                          # :158
            self.m[i] = self.m[i] * f(beta1) + f(1 - beta1) * g                              # :161
            G = self.m[i] / (np.sqrt(self.v[i]) + f(self.eps)) if n_sma > self.thr else self.m[i].copy()   # :187-191
            if self.wd != 0:
                G = G + f(self.wd) * self.p[i]                                               # :193-194
                if not n_sma > self.thr:
                    # reference quirk: below the threshold G_grad IS exp_avg (:191), so the in-place add_ of the weight
                    # decay term (:194) also lands in the stored first moment
                    self.m[i] = G.copy()
            self.p[i] = self.p[i] + f(-ss * self.lr) * G                                     # :199
            if self.step_no % self.k == 0:                                                   # :204-210
                self.slow[i] = self.slow[i] + f(self.alpha) * (self.p[i] - self.slow[i])
                self.p[i] = self.slow[i].copy()
        return out_g
