"""RAdam with the reference's constructor, update rule and state layout (radam.py:5-94).

CUDA parameters are updated by ONE fused kernel per contiguous run of parameters (pn_radam_step): the 16
hash tables are views of one [L,T,2] buffer and their gradients views of one flat buffer, so the whole
64 MiB table update is a single elementwise pass (7 x 64 MiB of traffic) instead of ~12 framework launches
per table.  The moments of such a run are allocated as one flat buffer and exposed per parameter, so
``state_dict()`` keeps the reference's per-parameter layout ('step', 'exp_avg', 'exp_avg_sq').  CPU
parameters (unit tests) use torch's multi-tensor ops."""
import math

import torch
from torch.optim.optimizer import Optimizer


class RAdam(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, degenerated_to_sgd=False):
        if not 0.0 <= lr:
            raise ValueError("Invalid learning rate: {}".format(lr))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {}".format(eps))
        if not (0.0 <= betas[0] < 1.0 and 0.0 <= betas[1] < 1.0):
            raise ValueError("Invalid betas: {}".format(betas))
        self.degenerated_to_sgd = degenerated_to_sgd
        # `buffer` is the reference's per-group memo of the rectification terms (radam.py:17-22, read at :60-78).  The
        # update here recomputes them (two pow() per step), but the key is kept in every param group so that a
        # checkpoint written by save_checkpoint loads into the reference's RAdam and vice versa.
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay,
                                      buffer=[[None, None, None] for _ in range(10)]))
        for group in self.param_groups:
            group["buffer"] = [[None, None, None] for _ in range(10)]

    @staticmethod
    def _rectification(step, beta1, beta2, degenerated_to_sgd):
        """radam.py:63-78 -> (N_sma, step_size)."""
        beta2_t = beta2 ** step
        n_max = 2 / (1 - beta2) - 1
        n_sma = n_max - 2 * step * beta2_t / (1 - beta2_t)
        if n_sma >= 5:
            step_size = math.sqrt((1 - beta2_t) * (n_sma - 4) / (n_max - 4) * (n_sma - 2) / n_sma * n_max / (n_max - 2)) \
                / (1 - beta1 ** step)
        elif degenerated_to_sgd:
            step_size = 1.0 / (1 - beta1 ** step)
        else:
            step_size = -1
        return n_sma, step_size

    @staticmethod
    def _runs(ps):
        """Split a list of parameters into maximal runs that are adjacent in memory (data and grad)."""
        runs, cur = [], []
        for p in ps:
            ok = (cur and p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                  and p.grad.dtype == torch.float32
                  and p.untyped_storage().data_ptr() == cur[-1].untyped_storage().data_ptr()
                  and p.grad.untyped_storage().data_ptr() == cur[-1].grad.untyped_storage().data_ptr()
                  and p.data_ptr() == cur[-1].data_ptr() + cur[-1].numel() * 4
                  and p.grad.data_ptr() == cur[-1].grad.data_ptr() + cur[-1].numel() * 4)
            if ok:
                cur.append(p)
            else:
                if cur:
                    runs.append(cur)
                cur = [p]
        if cur:
            runs.append(cur)
        return runs

    @staticmethod
    def _param_runs(ps):
        """Runs of parameters that are adjacent views of one storage (the hash tables)."""
        runs, cur = [], []
        for p in ps:
            ok = (cur and p.is_cuda and p.is_contiguous() and p.dtype == cur[-1].dtype
                  and p.untyped_storage().data_ptr() == cur[-1].untyped_storage().data_ptr()
                  and p.data_ptr() == cur[-1].data_ptr() + cur[-1].numel() * p.element_size())
            if ok:
                cur.append(p)
            else:
                if cur:
                    runs.append(cur)
                cur = [p]
        if cur:
            runs.append(cur)
        return runs

    def _init_state(self, ps):
        """Moments for parameters without state; adjacent parameters share one flat moment buffer."""
        new = [p for p in ps if len(self.state[p]) == 0]
        for run in self._param_runs(new):
            n = sum(p.numel() for p in run)
            m = torch.zeros(n, dtype=run[0].dtype, device=run[0].device)
            v = torch.zeros(n, dtype=run[0].dtype, device=run[0].device)
            off = 0
            for p in run:
                st = self.state[p]
                st["step"] = 0
                st["exp_avg"] = m[off:off + p.numel()].view_as(p)
                st["exp_avg_sq"] = v[off:off + p.numel()].view_as(p)
                off += p.numel()

    def _fused_cuda(self, ps, group, step, n_sma, step_size, dyn=None):
        from . import ops
        beta1, beta2 = group["betas"]
        mode = 2 if n_sma >= 5 else (1 if step_size > 0 else 0)
        lr, wd = group["lr"], group["weight_decay"]
        for run in self._runs(ps):
            first = run[0]
            n = sum(p.numel() for p in run)
            m0, v0 = self.state[first]["exp_avg"], self.state[first]["exp_avg_sq"]
            flat_state = all(self.state[p]["exp_avg"].data_ptr() == m0.data_ptr() + (p.data_ptr() - first.data_ptr())
                             and self.state[p]["exp_avg_sq"].data_ptr() == v0.data_ptr() + (p.data_ptr() - first.data_ptr())
                             and self.state[p]["exp_avg"].untyped_storage().data_ptr() == m0.untyped_storage().data_ptr()
                             and self.state[p]["exp_avg_sq"].untyped_storage().data_ptr() == v0.untyped_storage().data_ptr()
                             for p in run)
            if len(run) > 1 and flat_state:
                as1d = lambda t: torch.as_strided(t, (n,), (1,))
                args = [(as1d(first.data), as1d(first.grad), as1d(m0), as1d(v0))]
            else:
                args = [(p.data.view(-1), p.grad.view(-1), self.state[p]["exp_avg"].view(-1),
                         self.state[p]["exp_avg_sq"].view(-1)) for p in run]
            for pd, gd, m, v in args:
                if dyn is not None:
                    ops.radam_step_dyn(pd, gd, m, v, beta1, beta2, group["eps"], dyn, mode)
                else:
                    ops.radam_step(pd, gd, m, v, beta1, beta2, group["eps"], wd * lr, step_size * lr, mode)
            if getattr(self, "_recording", None) is not None:
                self._recording += args

    # ---- a step recorded in a CUDA graph (Trainer(cuda_graph=True)) ------------------------------------------------
    def _graph_groups(self):
        """[(group, live parameters)] when every parameter with a gradient can take the fused update in its rectified
        form (state initialised, N_sma >= 5 from here on), else None."""
        out = []
        for group in self.param_groups:
            live = [p for p in group["params"] if p.grad is not None]
            if not live:
                out.append((group, live))
                continue
            if not all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and not p.grad.is_sparse
                       and p.grad.is_contiguous() and p.grad.dtype == torch.float32 and len(self.state[p]) > 0
                       for p in live):
                return None
            steps = {self.state[p]["step"] for p in live}
            beta1, beta2 = group["betas"]
            if len(steps) != 1 or self._rectification(steps.pop() + 1, beta1, beta2, self.degenerated_to_sgd)[0] < 5:
                return None
            out.append((group, live))
        return out

    def graph_fingerprint(self, groups):
        """Every address a recorded update touches (parameters, gradients, both moments): a recorded iteration is only
        valid while none of them moved (load_state_dict, a re-flattened table set, a released gradient arena ...)."""
        fp = []
        for _, live in groups:
            for p in live:
                st = self.state[p]
                fp.append((p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()))
        return hash(tuple(fp))

    def graph_capture_step(self, dyn):
        """The launches of one update with the per-step scalars in dyn[group] = {wd*lr, step_size*lr} (device): called
        while the stream is capturing; touches no optimiser state."""
        for gi, (group, live) in enumerate(self._graph_groups()):
            if live:
                self._fused_cuda(live, group, None, 5, 1.0, dyn=dyn[gi])

    def graph_advance(self, groups=None):
        """Host side of one replay: advances every 'step' counter and returns the scalars of this update, flattened per
        group ([wd*lr, step_size*lr, ...]) as radam.py:63-85 computes them — or None when the update would not be the
        rectified one the graph recorded."""
        groups = groups if groups is not None else self._graph_groups()
        if groups is None:
            return None
        vals = []
        for group, live in groups:
            if not live:
                vals += [0.0, 0.0]
                continue
            beta1, beta2 = group["betas"]
            step = self.state[live[0]]["step"] + 1
            for p in live:
                self.state[p]["step"] = step
            _, step_size = self._rectification(step, beta1, beta2, self.degenerated_to_sgd)
            vals += [group["weight_decay"] * group["lr"], step_size * group["lr"]]
        return vals

    def _plan_key(self, live):
        st = self.state
        return tuple((p.data_ptr(), p.grad.data_ptr(), st[p]["exp_avg"].data_ptr() if p in st and "exp_avg" in st[p]
                      else 0) for p in live)

    def load_state_dict(self, state_dict):
        self._plans = {}
        return super().load_state_dict(state_dict)

    def _replay(self, gi, group, live):
        """The launches of the previous step again, when nothing moved (same parameters, same gradient buffers — the
        gradient arena keeps them fixed): skips the per-step re-derivation of the contiguous runs, which at a small
        per-GPU batch (strong scaling) was ~1 ms of host time per step."""
        plan = self._plans.get(gi)
        if plan is None or plan["key"] != self._plan_key(live):
            return False
        if any(self.state[p].get("step") != plan["step"] for p in live):
            return False
        from . import ops
        beta1, beta2 = group["betas"]
        step = plan["step"] + 1
        for p in live:
            self.state[p]["step"] = step
        n_sma, step_size = self._rectification(step, beta1, beta2, self.degenerated_to_sgd)
        mode = 2 if n_sma >= 5 else (1 if step_size > 0 else 0)
        lr, wd = group["lr"], group["weight_decay"]
        for pd, gd, m, v in plan["launches"]:
            ops.radam_step(pd, gd, m, v, beta1, beta2, group["eps"], wd * lr, step_size * lr, mode)
        plan["step"] = step
        return True

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        if not hasattr(self, "_plans"):
            self._plans = {}
        for gi, group in enumerate(self.param_groups):
            live0 = [p for p in group["params"] if p.grad is not None]
            if live0 and self._replay(gi, group, live0):
                continue
            self._recording = [] if all(p.is_cuda for p in live0) and live0 else None
            self._step_group(group)
            if self._recording is not None and self._recording:
                steps = {self.state[p]["step"] for p in live0}
                if len(steps) == 1:
                    self._plans[gi] = {"key": self._plan_key(live0), "step": steps.pop(), "launches": self._recording}
            self._recording = None
        return loss

    def _step_group(self, group):
        if True:
            beta1, beta2 = group["betas"]
            by_step = {}
            live = [p for p in group["params"] if p.grad is not None]
            for p in live:
                if p.grad.is_sparse:
                    raise RuntimeError("RAdam does not support sparse gradients")
            self._init_state(live)
            for p in live:
                st = self.state[p]
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            for step, ps in by_step.items():
                if all(p.is_cuda and p.dtype == torch.float32 and p.is_contiguous() and p.grad.is_contiguous()
                       and p.grad.dtype == torch.float32 for p in ps):
                    n_sma, step_size = self._rectification(step, beta1, beta2, self.degenerated_to_sgd)
                    self._fused_cuda(ps, group, step, n_sma, step_size)
                    continue
                grads = [p.grad for p in ps]
                m = [self.state[p]["exp_avg"] for p in ps]
                v = [self.state[p]["exp_avg_sq"] for p in ps]
                torch._foreach_mul_(v, beta2)
                torch._foreach_addcmul_(v, grads, grads, value=1 - beta2)          # radam.py:55
                torch._foreach_mul_(m, beta1)
                torch._foreach_add_(m, grads, alpha=1 - beta1)                     # radam.py:56
                n_sma, step_size = self._rectification(step, beta1, beta2, self.degenerated_to_sgd)
                lr, wd = group["lr"], group["weight_decay"]
                if n_sma >= 5:
                    if wd != 0:
                        torch._foreach_add_(ps, ps, alpha=-wd * lr)                # p += -wd*lr*p  (radam.py:82-83)
                    denom = torch._foreach_sqrt(v)
                    torch._foreach_add_(denom, group["eps"])
                    torch._foreach_addcdiv_(ps, m, denom, value=-step_size * lr)   # radam.py:84-85
                elif step_size > 0:
                    if wd != 0:
                        torch._foreach_add_(ps, ps, alpha=-wd * lr)
                    torch._foreach_add_(ps, m, alpha=-step_size * lr)

