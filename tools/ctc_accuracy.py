import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from oracle import ctc_oracle
from dae import _C
from dae.ctc import ctc_loss
from dae.greedy import greedy_ids_device
sys.path.insert(0, '/root/repo/tests')
from test_fullsize_gpu import _peaky_gpu
T, C = 2048, 4096
post = _peaky_gpu(T, C, C - 1, 300)[:, None]
_, ids, k = greedy_ids_device(post[:, 0], C - 1)
lab = ids[0, :int(k[0])].long()
tg = lab[None]
il = torch.tensor([T], device='cuda'); tl = torch.tensor([lab.numel()], device='cuda')
nll64, g64 = ctc_oracle.ctc_loss_grad(post.double().cpu().numpy(), tg.cpu().numpy(), [T], [lab.numel()], C - 1, gout=1.0 / T)
lp64 = post.double().cpu().numpy()
for path in ("0", "1"):
    _C.ctc_configure(blocked=int(path))
    x = post.clone().requires_grad_()
    nll = ctc_loss(x, tg, il, tl, blank=C - 1, reduction="none")
    (nll.sum() / T).backward()
    g = x.grad.cpu().numpy()
    tol = 1e-4 * np.abs(g64) + 1e-4 * (1.0 / T) * np.exp(lp64) + 1e-12
    r = np.abs(g - g64) / tol
    print("path", path, "loss rel err", abs(nll.item() - nll64[0]) / nll64[0], "max err / tol", r.max(), "at", np.unravel_index(r.argmax(), r.shape), "L", lab.numel())
y = post.clone().requires_grad_()
(torch.nn.functional.ctc_loss(y, tg, il, tl, blank=C - 1, reduction='sum') / T).backward()
r = np.abs(y.grad.cpu().numpy() - g64) / tol
print("torch cuda max err / tol", r.max())
