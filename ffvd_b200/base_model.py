"""Mirror of the hot-path part of `vfegpssm/base_model.py`: the adaptive SG-HMC update
(`generate_update_step`, :143-179), its schedule (`sghmc_step`, :915-933), `get_minibatch`
(:188-194) and the Adam step (`train_hypers`, :944-950).  Prediction / particle-Gibbs methods of
the reference class are out of scope."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np


class BaseModel(object):
    def __init__(self, Y, vars, minibatch_size, window_size):
        # base_model.py:12-27
        self.vars: List[str] = list(vars)
        self.minibatch_size = min(minibatch_size, self.X_N) if minibatch_size else self.X_N
        self.data_iter = 0
        self.window_size = window_size
        self.window: List[Dict[str, object]] = []
        self.posterior_samples = []
        self.sample_op = None
        self.burn_in_op = None
        self.adam_state: Dict[str, tuple] = {}
        self.adam_step = 0
        self.global_step = 1

    # ---- base_model.py:143-179
    def generate_update_step(self, nll=None, epsilon=0.01, mdecay=0.05):
        """Allocate the SG-HMC state (xi = g = g2 = 1, p = 0, :151-154) for every sampled variable.
        `nll` is accepted for signature parity (the reference passes its graph tensor)."""
        import torch
        self.epsilon = epsilon
        self.mdecay = mdecay
        self.sghmc_state = {}
        for name in self.vars:
            th = self.params[name]
            self.sghmc_state[name] = dict(xi=torch.ones_like(th), g=torch.ones_like(th), g2=torch.ones_like(th),
                                          p=torch.zeros_like(th))
        self.burn_in_op = "burn_in"
        self.sample_op = "sample"

    def _run_update(self, burn_in: bool, noise: Optional[Dict[str, object]] = None):
        """One `session.run(burn_in_op | sample_op)`: a full nll+gradient evaluation followed by the
        elementwise update of every sampled variable (Jacobi semantics)."""
        import torch
        out = self.evaluate()
        for name in self.vars:
            th = self.params[name]
            st = self.sghmc_state[name]
            nz = noise[name] if noise is not None else torch.randn(th.shape, dtype=torch.float64, device=th.device)
            nz = torch.as_tensor(np.asarray(nz), dtype=torch.float64, device=th.device) if not torch.is_tensor(nz) else nz
            self.ctx.sghmc_update(th, out["g_" + name], nz.contiguous(), st["xi"], st["g"], st["g2"], st["p"],
                                  self.epsilon, self.mdecay, float(self.X_N), burn_in)
        return out

    def get_minibatch(self, global_step=1):
        # base_model.py:188-194: always the full batch; decayed Adam learning rate
        return [0, self.X_N], 0.003 * (0.95 ** (global_step / 1000))

    # ---- base_model.py:915-933
    def sghmc_step(self, noise_fn=None):
        """1 burn-in update, then 10 x (burn-in, sample); snapshot the sampled variables into the
        window (21 nll+gradient evaluations)."""
        k = 0

        def nz():
            nonlocal k
            k += 1
            return None if noise_fn is None else noise_fn(k - 1)
        self._run_update(True, nz())
        for _ in range(10):
            self._run_update(True, nz())
            self._run_update(False, nz())
        sample = {name: self.params[name].clone() for name in self.vars}
        self.window.append(sample)
        if len(self.window) > self.window_size:
            self.window = self.window[-self.window_size:]

    # ---- base_model.py:944-950 + dgp_model.py:303-305
    def train_hypers(self, window_index: Optional[int] = None):
        """One TF1-Adam step on the trainable set, with the SG-HMC variables fed from a random
        window entry (the feed is temporary, as in the reference's feed_dict)."""
        import torch
        _, lr = self.get_minibatch(self.global_step)
        saved = {}
        if self.window:
            i = np.random.randint(len(self.window)) if window_index is None else window_index
            for name, val in self.window[i].items():
                saved[name] = self.params[name].clone()
                self.params[name].copy_(val)
        out = self.evaluate()
        for name, val in saved.items():
            self.params[name].copy_(val)
        self.adam_step += 1
        for name in self.trainable:
            th = self.params[name]
            if name not in self.adam_state:
                self.adam_state[name] = (torch.zeros_like(th), torch.zeros_like(th))
            m, v = self.adam_state[name]
            self.ctx.adam_update(th, out["g_" + name], m, v, lr, 0.9, 0.999, 1e-8, self.adam_step)
        return out
