"""`fossen.bluerov_torch` mirror: reduced 9-state RHS on the CUDA engine (reference: fossen/bluerov_torch.py:8-67)."""
import math

import torch

from ..engine import reduced9_rhs


def ssa(angle):
    """Smallest signed angle, `angle - 2 pi floor((angle + pi) / (2 pi))` (fossen/bluerov_torch.py:8-18)."""
    return angle - 2 * math.pi * torch.floor_divide(angle + math.pi, 2 * math.pi)


def bluerov_compute(t, x_, u_):
    """x_dot [B,9] of the reduced model for x_ [B,9] = [x, y, z, cos psi, sin psi, u, v, w, r] and u_ [B,4] =
    [X, Y, Z, M_z]; 1-D arguments are promoted to a batch of one (fossen/bluerov_torch.py:20-67).  `t` is unused, as
    in the reference.  The result has the dtype and device of x_; the arithmetic always runs on the GPU."""
    x = x_.unsqueeze(0) if x_.dim() == 1 else x_
    u = u_.unsqueeze(0) if u_.dim() == 1 else u_
    if x.dim() != 2 or x.shape[1] != 9 or u.dim() != 2 or u.shape[1] != 4 or u.shape[0] != x.shape[0]:
        raise ValueError(f"expected x_ [B,9] and u_ [B,4], got {tuple(x_.shape)} and {tuple(u_.shape)}")
    dev = x.device
    work = torch.promote_types(x.dtype, u.dtype)
    if work not in (torch.float32, torch.float64):
        work = torch.float32
    cuda = dev if dev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
    out = reduced9_rhs(x.detach().to(cuda, work), u.detach().to(cuda, work))
    return out.to(device=dev, dtype=x.dtype)
