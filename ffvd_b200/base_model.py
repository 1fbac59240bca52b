"""Mirror of the hot-path part of `vfegpssm/base_model.py`: the adaptive SG-HMC update
(`generate_update_step`, :143-179), its schedule (`sghmc_step`, :915-933), `get_minibatch`
(:188-194), the Adam step (`train_hypers`, :944-950) and the posterior roll-out / results file of
`collect_samples_formal` (:197-522) and the particle-Gibbs sweep for X (`PG_for_X`, :29-75)."""
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
        from . import _capi
        if not self.vars:
            # cases 1, 4, 6: nothing is SG-HMC sampled, burn_in_op / sample_op are empty lists and session.run([]) evaluates
            # nothing -- no nll evaluation here either
            return None
        # Z and the kernel hyper-parameters are fixed across the 21 evaluations of a step unless they are sampled: the
        # Cholesky factors of K(Z,Z) are then reused (FFVD_FLAG_REUSE_KZZ)
        out = self.evaluate(_capi.FLAG_REUSE_KZZ if getattr(self, "_kzz_clean", False) else 0)
        self._kzz_clean = not any(n in ("Z", "logv", "logl") for n in self.vars)
        for name in self.vars:
            th = self.params[name]
            st = self.sghmc_state[name]
            nz = noise[name] if noise is not None else torch.randn(th.shape, dtype=torch.float64, device=th.device)
            nz = torch.as_tensor(np.asarray(nz), dtype=torch.float64, device=th.device) if not torch.is_tensor(nz) else nz
            self.ctx.sghmc_update(th, out["g_" + name], nz.contiguous(), st["xi"], st["g"], st["g2"], st["p"],
                                  self.epsilon, self.mdecay, float(self.X_N), burn_in)
        return out

    def invalidate_kzz(self):
        """Call after changing Z or the kernel hyper-parameters from outside this class (the SG-HMC loop otherwise
        re-uses the Cholesky factors of K(Z,Z) between evaluations that cannot have moved them)."""
        self._kzz_clean = False

    def get_minibatch(self, global_step=1):
        # base_model.py:188-194: always the full batch; decayed Adam learning rate
        return [0, self.X_N], 0.003 * (0.95 ** (global_step / 1000))

    def enable_graph(self, enable: bool = True):
        """Run `sghmc_step` as ONE CUDA-graph launch: the reference's 21 `session.run` calls (base_model.py:919-925) are 21
        evaluations x ~10 kernels + 21 x len(vars) updates; captured once, replayed per step.  The injected noise lives in
        fixed buffers refilled before every launch (graphs capture addresses, not values)."""
        self._use_graph = bool(enable)
        g = getattr(self, "_graph", None)
        if g is not None:
            g.close()
        self._graph = None

    def _sghmc_step_graph(self, noise_fn=None):
        import torch
        from . import _capi
        ctx = self.ctx
        if getattr(self, "_graph", None) is None:
            self.evaluate()                               # lays out the workspace and the output tensors (capture needs both)
            self._gnoise = {n: torch.empty((21,) + tuple(self.params[n].shape), dtype=torch.float64, device=self.device) for n in self.vars}
            prob = dict(self.params); prob.update(self.data); prob.setdefault("logl", None)
            fixed_kzz = not any(n in ("Z", "logv", "logl") for n in self.vars)
            X_N = float(self.X_N)

            def body():
                for k in range(21):
                    burn_in = (k == 0) or (k % 2 == 1)    # burn-in, then 10 x (burn-in, sample): base_model.py:919-925
                    fl = self.flags | _capi.FLAG_ASYNC | (_capi.FLAG_REUSE_KZZ if (k > 0 and fixed_kzz) else 0)
                    ctx.nll_grads(self.kind, self.U_collapse, prob, self._out, flags=fl, jitter=1e-5)
                    for n in self.vars:
                        st = self.sghmc_state[n]
                        ctx.sghmc_update(self.params[n], self._out["g_" + n], self._gnoise[n][k], st["xi"], st["g"], st["g2"], st["p"],
                                         self.epsilon, self.mdecay, X_N, burn_in)
            self._graph = ctx.capture(body, keep=(prob, self._out, self._gnoise, self.sghmc_state))
        for n in self.vars:
            if noise_fn is None:
                self._gnoise[n].normal_()
            else:
                for k in range(21):
                    nz = noise_fn(k)[n]
                    self._gnoise[n][k].copy_(nz if torch.is_tensor(nz) else torch.as_tensor(np.asarray(nz), dtype=torch.float64, device=self.device))
        self._graph.launch()
        self._kzz_clean = False
        if not bool(torch.isfinite(self._out["nll"]).all()):     # no status read-back inside a graph: NaN is the failure signal
            raise FloatingPointError("sghmc_step (CUDA graph): non-finite nll -- K(Z,Z) + jitter not positive definite, or a diverged chain")

    # ---- base_model.py:915-933
    def sghmc_step(self, noise_fn=None):
        """1 burn-in update, then 10 x (burn-in, sample); snapshot the sampled variables into the
        window (21 nll+gradient evaluations)."""
        if getattr(self, "_use_graph", False) and self.vars:
            self._sghmc_step_graph(noise_fn)
            sample = {name: self.params[name].clone() for name in self.vars}
            self.window.append(sample)
            if len(self.window) > self.window_size:
                self.window = self.window[-self.window_size:]
            return
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
        # base_model.py:945: get_minibatch() is called WITHOUT an argument, so global_step is always 1 and the Adam rate is
        # the constant 0.003 * 0.95**(1/1000); set `model.decay_lr = True` for the decaying schedule the formula
        # suggests (an opt-in, not the reference's behaviour)
        _, lr = self.get_minibatch(self.global_step) if getattr(self, "decay_lr", False) else self.get_minibatch()
        saved = {}
        self._kzz_clean = False                       # the window feed and the Adam step may move Z / hyper-parameters
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

    # ---- base_model.py:197-522
    def collect_samples_formal(self, num, spacing, control_inputs, test_len, sghmc_var_len=0, U_collapse=False, Y_test=None,
                               Y_train_std=1., save_path_file=None, Y_train=None, case='C1', ll_seq=[0.], running_time_seq=[0.],
                               PG_num=None, synthetic_data_function_plot=False, data_uu=None, noise=None):
        """Posterior roll-out, `base_model.py:197-522`: for each of `num` posterior samples (advanced by `spacing`
        SG-HMC sample updates when anything is sampled, :226-232), start from the last latent state and iterate
        x_{t+1} = x_t + f(x_t, c_t) + eps sqrt(var + Q) for `test_len` steps (:283-310) with the prediction-time
        conditional (q(u) = the collapsed optimum when `U_collapse`, :241-252); then y = x C + d, RMSE over the
        first 30 test steps (:330-347) and the `_results.npz` file with the reference's keys (:488-517).
        `noise` (num, test_len, D) replaces the reference's unseeded tf.random.normal draws.
        When nothing is SG-HMC sampled every posterior sample shares its parameters and the `num` roll-outs advance
        together as one N = num conditional per step."""
        import torch
        from . import conditionals_multi_output as cmo
        if synthetic_data_function_plot:
            raise NotImplementedError("synthetic_data_function_plot is a plotting aid of the reference (base_model.py:262-278)")
        dev = self.device
        t64 = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
        X = self.params["X"]
        if X.dim() != 2:
            raise NotImplementedError("roll-out of a batched-sample model: pick one trajectory")
        D = X.shape[1]
        kern = self.kernels[-1]
        ctrl_all = t64(control_inputs) if control_inputs is not None else torch.zeros((0, 0), dtype=torch.float64, device=dev)
        n_ctrl = ctrl_all.shape[1] if ctrl_all.dim() == 2 else 0
        T_train = int(np.asarray(Y_train).shape[0]) if Y_train is not None else self.X_N - 1
        L = int(test_len)                                           # prediction_length = test_len + pre_index - 1, pre_index = 1
        if noise is None:
            noise = torch.randn((num, L, D), dtype=torch.float64, device=dev)
        else:
            noise = t64(noise)
        Qv = self.log_Q.exp()
        self.fit_x = X.clone()

        def posterior_qu():
            if U_collapse:
                xc = torch.cat((X[:self.X_N - 1], ctrl_all[:self.X_N - 1]), dim=1) if n_ctrl > 0 else X[:-1]
                U_val, Lseq = cmo.collapse_u_mean_after_kernel_precalculation(None, xc.contiguous(), X, self.params["Z"], kern, Qv)
                return U_val[0], Lseq
            return self.params["U"], None

        def roll(x_t, U_val, q_sqrt, eps):
            """x_t (n,D), eps (n,L,D) -> states (n,L,D), variances (n,L,D)"""
            xs, vs = [], []
            # pre-calculate the Cholesky decomposition of the kernel matrix (:209, :235): one factorisation per roll-out,
            # reused by the L per-step conditionals
            factors = cmo.kernel_pre_cal(self.params["Z"], kern)
            for t in range(L):
                if n_ctrl > 0:
                    c = ctrl_all[t + T_train][None, :].expand(x_t.shape[0], n_ctrl)
                    xc = torch.cat((x_t, c), dim=1).contiguous()
                else:
                    xc = x_t.contiguous()
                mu, var = cmo.conditional_after_kernel_precalculation(factors, xc, self.params["Z"], kern, U_val, white=True,
                                                                      full_cov=False, q_sqrt=q_sqrt)
                x_next = mu + x_t + eps[:, t] * torch.sqrt(var + Qv)
                xs.append(x_next)
                vs.append(var + Qv)
                x_t = x_next
            return torch.stack(xs, dim=1), torch.stack(vs, dim=1)

        mc_posterior_samples = [[] for _ in range(sghmc_var_len)]
        if sghmc_var_len == 0:
            U_val, q_sqrt = posterior_qu()
            x0 = X[-1][None, :].expand(num, D).contiguous()
            predict_x_whole, predict_x_var_whole = roll(x0, U_val, q_sqrt, noise)
        else:
            px, pv = [], []
            for i in range(num):
                for _ in range(spacing):
                    self._run_update(False)
                for j, name in enumerate(self.vars[:sghmc_var_len]):
                    mc_posterior_samples[j].append(self.params[name].cpu().numpy().copy())
                U_val, q_sqrt = posterior_qu()
                a, b = roll(X[-1][None, :].contiguous(), U_val, q_sqrt, noise[i:i + 1])
                px.append(a[0]); pv.append(b[0])
            predict_x_whole, predict_x_var_whole = torch.stack(px), torch.stack(pv)

        predict_x_whole = predict_x_whole.cpu().numpy()
        predict_x_var_whole = predict_x_var_whole.cpu().numpy()
        CC_val = self.params["C"].cpu().numpy()
        DD_val = self.params["d"].cpu().numpy()
        log_R_cholesky = self.params["logR"].cpu().numpy()
        fit_x_value = self.fit_x.cpu().numpy()[1:]
        self.predict_y = (np.mean(np.einsum('ijk,kl->ijl', predict_x_whole, CC_val), axis=0) + DD_val[None, :]).reshape((-1))
        self.predict_y_var = (np.mean(np.einsum('ijk,kl->ijl', predict_x_var_whole, CC_val ** 2), axis=0)).reshape((-1)) \
            + np.exp(2 * log_R_cholesky).reshape(-1)
        self.fit_y = (np.matmul(fit_x_value, CC_val) + DD_val).reshape((-1))
        self.predict_x, self.predict_x_var = predict_x_whole, predict_x_var_whole
        self.RMSE_val = None
        if Y_test is not None:
            Y_test = np.asarray(Y_test)
            n30 = min(30, self.predict_y.shape[0])
            self.RMSE_val = float(np.sqrt(np.mean((Y_test[:30].reshape((-1))[:n30] - self.predict_y[:n30]) ** 2)) * Y_train_std)
        if save_path_file is not None:
            k_log_lengthscales = [np.asarray(k.loglengthscales.cpu().numpy() if hasattr(k.loglengthscales, "cpu") else k.loglengthscales)
                                  for k in kern] if self.kind == 0 else []
            k_log_variances = [np.asarray(k.logvariance.cpu().numpy() if hasattr(k.logvariance, "cpu") else k.logvariance) for k in kern]
            np.savez_compressed(save_path_file + '_results.npz', y_train_vfe=self.fit_y, y_test_vfe=self.predict_y,
                                v_test_vfe_var=self.predict_y_var, Y_test_data=Y_test, Y_train_data=Y_train, Y_train_std=Y_train_std,
                                CC_val=CC_val, DD_val=DD_val, log_R_cholesky=log_R_cholesky, log_QQ=self.log_Q.cpu().numpy(),
                                Z_val=self.params["Z"].cpu().numpy(), U_val=self.params["U"].cpu().numpy(), X_val=fit_x_value,
                                k_lengthscales=k_log_lengthscales, k_log_variances=k_log_variances, case=case, ll_seq=ll_seq,
                                running_time_seq=running_time_seq, PG_num=PG_num,
                                mc_posterior_samples=np.asarray(mc_posterior_samples, dtype=object) if sghmc_var_len else [])
        return self.predict_y, self.predict_y_var

    # ---- base_model.py:29-75 / 78-138
    def PG_for_X(self, control_inputs, PG_particles, normals=None, eps=None, uniforms=None, assign=True):
        """Conditional SMC (particle Gibbs) sweep over the latent trajectory, `base_model.py:29-75`: PG_particles - 1 free
        particles plus the current trajectory as the reference particle; per time step one N = P-1 prediction-time
        conditional (factors of K(Z,Z) computed once, `kernel_pre_cal`), Gaussian emission weights, multinomial resampling;
        the returned / assigned trajectory is drawn from the final weights.  Trajectories are kept as per-step states and
        ancestor indices and traced back at the end (the reference gathers whole trajectories at every step: same result).
        `normals` (P-1,D), `eps` (T,P-1,D), `uniforms` (T,P-1) inject the randomness (the reference draws unseeded).
        NB the reference's graph-mode `PG_for_X_speedup` (:78-138), the variant its driver wires up, never executes its
        assign (the op returned to `session.run` is `tf.ones(1)`) and drops its TensorArray writes: what is mirrored here
        is the algorithm those two methods describe."""
        import torch
        from . import conditionals_multi_output as cmo
        dev = self.device
        t64 = lambda a: a.to(dtype=torch.float64, device=dev) if torch.is_tensor(a) else torch.as_tensor(np.asarray(a, dtype=np.float64), device=dev)
        X = self.params["X"]
        if X.dim() != 2:
            raise NotImplementedError("PG_for_X works on one trajectory")
        T, D = X.shape[0] - 1, X.shape[1]
        P1 = int(PG_particles) - 1
        ctrl = t64(control_inputs) if control_inputs is not None else torch.zeros((T, 0), dtype=torch.float64, device=dev)
        n_ctrl = ctrl.shape[1] if ctrl.dim() == 2 else 0
        normals = torch.randn((P1, D), dtype=torch.float64, device=dev) if normals is None else t64(normals)
        eps = torch.randn((T, P1, D), dtype=torch.float64, device=dev) if eps is None else t64(eps)
        uniforms = torch.rand((T, P1), dtype=torch.float64, device=dev) if uniforms is None else t64(uniforms)
        kern = self.kernels[-1]
        from .likelihoods import logdensity_norm
        Z, U = self.params["Z"], self.params["U"]
        Rchols = self.likelihood.Rchols                       # (Dy,Dy) lower factor; likelihoods.py:114-127 on the device
        Qv = self.log_Q.exp()
        Y = self.data["Y"]
        factors = cmo.kernel_pre_cal(Z, kern)                  # once per sweep (:36)
        states = torch.empty((T + 1, P1 + 1, D), dtype=torch.float64, device=dev)
        states[0, :P1] = normals
        states[:, P1] = X                                      # reference particle
        anc = torch.empty((max(T - 1, 0), P1), dtype=torch.long, device=dev)
        x_t = normals

        def sample(logits, u):
            cdf = torch.cumsum(torch.softmax(logits, dim=0), dim=0)
            return torch.clamp(torch.searchsorted(cdf, u.reshape(-1).contiguous()), max=logits.shape[0] - 1)

        last = None
        for tt in range(T):
            xc = torch.cat((x_t, ctrl[tt][None, :].expand(P1, n_ctrl)), dim=1).contiguous() if n_ctrl > 0 else x_t.contiguous()
            mu, var = cmo.conditional_after_kernel_precalculation(factors, xc, Z, kern, U, white=True, full_cov=False)
            x_next = mu + x_t + eps[tt] * torch.sqrt(var + Qv)
            states[tt + 1, :P1] = x_next
            y_mu = self.likelihood.predict_mean(states[tt + 1].contiguous())   # (P, Dy): free particles and the reference one (:61-65)
            logits = logdensity_norm(Y[tt].contiguous(), y_mu, Rchols)
            if tt < T - 1:
                idx = sample(logits, uniforms[tt])
                anc[tt] = idx
                x_t = states[tt + 1][idx]
            else:
                last = sample(logits, uniforms[tt][:1])[0]
        # trace the chosen particle back through its ancestors
        path = torch.empty((T + 1, D), dtype=torch.float64, device=dev)
        k = last
        path[T] = states[T, k]
        for tt in range(T - 1, -1, -1):
            # slot k at time tt+1 (free slots only) descends from anc[tt-1][k] at time tt; the reference slot from itself
            if tt >= 1:
                k = torch.where(k < P1, anc[tt - 1][torch.clamp(k, max=P1 - 1)], k)
            path[tt] = states[tt, k]
        if assign:
            X.copy_(path)
        return path

    def gp_x_sampling(self):
        """`base_model.py:936-942`."""
        return self.PG_for_X(self.data.get("ctrl"), getattr(self, "PG_particles", 100))
