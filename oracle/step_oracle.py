"""CPU restatement (numpy, float32 op by op) of the elementwise work either side of the rasterizer in a training step.
*** TEST INFRASTRUCTURE ONLY *** — imported by tests/ only; the product package never imports it.

  activate / activate_backward : fov3dgs/scene/gaussian_model.py:40-60 (scaling_activation = torch.exp, rotation_activation =
      torch.nn.functional.normalize, opacity_activation = torch.sigmoid) and their autograd derivatives
  adam_step : the optimizer of scene/gaussian_model.py:289, torch.optim.Adam(l, lr=0.0, eps=1e-15), stepped at
      eff_finetune.py:146.  torch is a third-party dependency of the reference (not vendored under /root/reference); the
      algorithm restated here is torch/optim/adam.py `_single_tensor_adam` of the installed torch 2.11:
          exp_avg.lerp_(grad, 1 - beta1); exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
          denom = exp_avg_sq.sqrt() / sqrt(1 - beta2**t) + eps;  param.addcdiv_(exp_avg, denom, value=-lr / (1 - beta1**t))
Pinned by tests/test_step_oracle.py against torch.optim.Adam / torch autograd on CPU (the same dependency the reference runs).
"""
import numpy as np

F = np.float32


def activate(raw_scale, raw_rot, raw_opacity):
    q = raw_rot.astype(F)
    n = np.sqrt((q * q).sum(axis=1, keepdims=True, dtype=F))
    d = np.maximum(n, F(1e-12))
    return np.exp(raw_scale.astype(F)), q / d, F(1) / (F(1) + np.exp(-raw_opacity.astype(F)))


def activate_backward(raw_rot, scale, opacity, d_scale, d_rot, d_opacity):
    q = raw_rot.astype(np.float64)
    g = d_rot.astype(np.float64)
    n = np.sqrt((q * q).sum(axis=1, keepdims=True))
    dn = np.maximum(n, 1e-12)
    ok = (n >= 1e-12) & (n > 0)
    k = np.where(ok, (g * q).sum(axis=1, keepdims=True) / np.where(ok, dn * dn * n, 1.0), 0.0)
    d_raw_rot = g / dn - q * k
    return (d_scale * scale).astype(F), d_raw_rot.astype(F), (d_opacity * (F(1) - opacity) * opacity).astype(F)


def adam_step(param, grad, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place on the float32 arrays; `step` is the 1-based update count.  Scalars are Python doubles narrowed once."""
    w1, b2, w2 = F(1 - beta1), F(beta2), F(1 - beta2)
    bc1, bc2 = 1 - beta1 ** step, 1 - beta2 ** step
    neg_step, bc2_sqrt, e = F(-(lr / bc1)), F(bc2 ** 0.5), F(eps)
    g = grad.astype(F)
    diff = g - exp_avg
    if abs(w1) < 0.5:
        exp_avg += w1 * diff
    else:
        exp_avg[...] = g - diff * (F(1) - w1)
    exp_avg_sq *= b2
    exp_avg_sq += w2 * (g * g)
    denom = np.sqrt(exp_avg_sq) / bc2_sqrt + e
    param += neg_step * (exp_avg / denom)
