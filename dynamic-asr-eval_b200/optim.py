"""MADGRAD (Defazio & Jelassi 2021), the optimizer the reference adapts with
(``lcasr.optim.madgrad.MADGRAD``, un-vendored; used as the default ``optim=`` at lcasr/lib.py:458
with ``lr=9e-5``).  Restated from the published algorithm with torch foreach ops; defaults
momentum=0.9, weight_decay=0, eps=1e-6 as in the facebookresearch/madgrad release.
"""
import torch


class MADGRAD(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-2, momentum=0.9, weight_decay=0.0, eps=1e-6):
        if not 0 <= momentum < 1:
            raise ValueError("momentum must be in [0,1)")
        super().__init__(params, dict(lr=lr, momentum=momentum, weight_decay=weight_decay, eps=eps))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            lr, mom, wd, eps = group["lr"] + group["eps"], group["momentum"], group["weight_decay"], group["eps"]
            ck = 1.0 - mom
            for p in ps:
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["grad_sum_sq"] = torch.zeros_like(p)
                    st["s"] = torch.zeros_like(p)
                    if mom != 0:
                        st["x0"] = p.detach().clone()
            k = self.state[ps[0]]["step"]
            lamb = lr * (k + 1) ** 0.5
            grads = [p.grad for p in ps]
            if wd != 0:
                grads = torch._foreach_add(grads, ps, alpha=wd)
            gss = [self.state[p]["grad_sum_sq"] for p in ps]
            ss = [self.state[p]["s"] for p in ps]
            if mom == 0:
                rms = torch._foreach_pow(gss, 1.0 / 3)
                torch._foreach_add_(rms, eps)
                x0 = torch._foreach_addcdiv(ps, ss, rms, value=1.0)
            else:
                x0 = [self.state[p]["x0"] for p in ps]
            torch._foreach_addcmul_(gss, grads, grads, value=lamb)
            rms = torch._foreach_pow(gss, 1.0 / 3)
            torch._foreach_add_(rms, eps)
            torch._foreach_add_(ss, grads, alpha=lamb)
            z = torch._foreach_addcdiv(x0, ss, rms, value=-1.0)
            if mom == 0:
                torch._foreach_copy_(ps, z)
            else:
                torch._foreach_mul_(ps, 1.0 - ck)
                torch._foreach_add_(ps, z, alpha=ck)
            for p in ps:
                self.state[p]["step"] = k + 1
        return loss
