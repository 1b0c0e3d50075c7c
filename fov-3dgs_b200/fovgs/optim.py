"""Adam for the Gaussian model as one multi-tensor launch per step (include/fovgs.h: fovgs_adam_step).

Drop-in for the optimizer the reference builds in scene/gaussian_model.py:279-289,
`torch.optim.Adam(l, lr=0.0, eps=1e-15)` with one parameter group per model tensor, stepped at eff_finetune.py:146.
The state layout is torch.optim.Adam's (`state[p] = {"step", "exp_avg", "exp_avg_sq"}`), so the reference's optimizer surgery
(replace_tensor_to_optimizer / _prune_optimizer / cat_tensors_to_optimizer, gaussian_model.py:609-696) and
state_dict()/load_state_dict() keep working on it.  There is no CPU path: parameters must be float32 CUDA tensors.
"""
import ctypes as C

import torch

from ._lib import ADAM_MAX_GROUPS, AdamGroup, check, lib


class Adam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False):
        if weight_decay != 0 or amsgrad or maximize:
            raise NotImplementedError("fovgs.optim.Adam implements the reference's configuration only: weight_decay=0, "
                                      "amsgrad=False, maximize=False")
        if lr < 0.0:
            raise ValueError(f"Invalid learning rate: {lr}")
        if eps < 0.0:
            raise ValueError(f"Invalid epsilon value: {eps}")
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 0: {betas[0]}")
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError(f"Invalid beta parameter at index 1: {betas[1]}")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False, maximize=False))

    def _collect(self):
        """One fovgs_adam_group per parameter that has a gradient; lazily creates the state like torch does."""
        out, keep = [], []
        for group in self.param_groups:
            b1, b2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if not p.is_cuda or p.dtype != torch.float32:
                    raise RuntimeError("fovgs.optim.Adam: parameters must be float32 CUDA tensors (no CPU fallback)")
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients")
                if not p.is_contiguous():
                    raise RuntimeError("fovgs.optim.Adam: parameters must be contiguous")
                st = self.state[p]
                if len(st) == 0:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                for k in ("exp_avg", "exp_avg_sq"):
                    if not st[k].is_contiguous() or st[k].shape != p.shape or st[k].device != p.device:
                        raise RuntimeError(f"fovgs.optim.Adam: state '{k}' must be a contiguous tensor of the parameter's shape")
                st["step"] += 1
                g = p.grad if p.grad.is_contiguous() else p.grad.contiguous()
                if g.dtype != torch.float32:
                    g = g.float()
                keep.append(g)
                a = AdamGroup()
                a.param, a.grad, a.exp_avg, a.exp_avg_sq = p.data_ptr(), g.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr()
                a.n, a.step = p.numel(), int(st["step"].item())
                a.lr, a.beta1, a.beta2, a.eps = float(group["lr"]), float(b1), float(b2), float(group["eps"])
                out.append((p.device, a))
        return out, keep

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        groups, keep = self._collect()
        by_dev = {}
        for dev, a in groups:
            by_dev.setdefault(dev, []).append(a)
        for dev, items in by_dev.items():
            stream = torch.cuda.current_stream(dev).cuda_stream
            with torch.cuda.device(dev):
                for i in range(0, len(items), ADAM_MAX_GROUPS):
                    chunk = items[i:i + ADAM_MAX_GROUPS]
                    arr = (AdamGroup * len(chunk))(*chunk)
                    check(lib().fovgs_adam_step(arr, len(chunk), C.c_void_p(stream)), "fovgs_adam_step")
        del keep
        return loss
