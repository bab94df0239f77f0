"""RAdam with the reference's constructor, update rule and state layout (radam.py:5-94), written with
torch's multi-tensor (_foreach) ops: one fused launch per elementwise stage over all parameters of a
group instead of ~12 launches per parameter.  The optimiser stays PyTorch by design (BASELINE.json
north_star); a flat fused CUDA step is the first item of SURVEY.md §8f."""
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
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))

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

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            by_step = {}
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("RAdam does not support sparse gradients")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                by_step.setdefault(st["step"], []).append(p)
            for step, ps in by_step.items():
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
        return loss
