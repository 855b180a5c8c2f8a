"""numpy restatement of the engine's in-kernel command-signal generator — TEST INFRASTRUCTURE (see oracle/__init__.py):
only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import it.

The signal is the reference's "random but smooth thruster command" (training/train_sim_brov2_koopmanEDMDc.py:161-164,180)

    u_k = clip(alpha u_{k-1} + sigma N(0,1), -1, 1),   alpha = 0.98, sigma = 0.02, u_{-1} = 0

with the normal deviates drawn from a counter-based generator instead of numpy's global Mersenne twister:
Philox4x32-10 (J. Salmon, M. Moraes, R. Dror, D. Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11; the
Random123 library's philox4x32 with 10 rounds), key = the 64-bit seed, counter = (vehicle lo, vehicle hi, step lo,
(step hi << 1) | block) with block 0 -> channels 0..3, block 1 -> channels 4..7; each pair of 32-bit words (a, b) gives
two deviates by Box-Muller on 23-bit uniforms.  The integer part is bit-exact with the device code; the transcendental
part uses numpy's float32 log/sin/cos where the kernel uses the GPU's fast-math units, so the deviates agree to ~1e-6
— which is why parity tests feed the oracle the engine's own materialised inputs (brov_generate_inputs) and use this
module only to pin the generator itself.
"""
from __future__ import annotations

import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: uint32 array [..., 4]; key: (k0, k1) python ints -> uint32 array [..., 4]."""
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def box_muller(a, b):
    """Two float32 N(0,1) deviates from two uint32 words, as the device code forms them."""
    f1 = ((a >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)
    f2 = ((b >> np.uint32(9)) | np.uint32(0x3F800000)).view(np.float32)
    u1 = f1 - np.float32(0.99999994)
    th = f2 * np.float32(6.2831855) + np.float32(-6.2831855)
    r = np.sqrt(np.float32(-1.3862944) * np.log2(u1, dtype=np.float32), dtype=np.float32)
    return r * np.cos(th, dtype=np.float32), r * np.sin(th, dtype=np.float32)


def normals(seed: int, vehicles, steps):
    """N(0,1) deviates [len(steps), len(vehicles), 8] (float32) of the given global vehicle / step indices."""
    v = np.asarray(vehicles, dtype=np.uint64)[None, :]
    s = np.asarray(steps, dtype=np.uint64)[:, None]
    v, s = np.broadcast_arrays(v, s)
    key = (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    out = np.zeros(v.shape + (8,), np.float32)
    for blk in (0, 1):
        ctr = np.stack([v & MASK, v >> np.uint64(32), s & MASK, ((s >> np.uint64(32)) << np.uint64(1)) | np.uint64(blk)],
                       axis=-1).astype(np.uint32)
        w = philox4x32_10(ctr, key)
        for pair in (0, 1):
            n0, n1 = box_muller(w[..., 2 * pair], w[..., 2 * pair + 1])
            out[..., 4 * blk + 2 * pair], out[..., 4 * blk + 2 * pair + 1] = n0, n1
    return out


def command_signal(seed: int, vehicles, step0: int, steps: int, nu: int = 8, rho=0.98, sigma=0.02, clip=1.0, scale=None,
                   state0=None, dtype=np.float64):
    """U [steps, len(vehicles), nu] and the AR(1) state after the last step.  As in the engine the recursion
    s <- clip(rho s + (sigma scale_j) n, -(clip scale_j), clip scale_j) runs in float32 whatever the engine's scalar
    type (the signal is a sequence of float32-representable numbers); the result is returned as `dtype`."""
    veh = np.asarray(vehicles)
    f = np.float32
    n = normals(seed, veh, np.arange(step0, step0 + steps))[..., :nu]
    sc = np.ones(nu) if scale is None else np.asarray(scale, float)
    sg, cl = (sigma * sc).astype(f), (clip * sc).astype(f)
    s = np.zeros((len(veh), nu), f) if state0 is None else np.array(state0, f)
    U = np.zeros((steps, len(veh), nu), f)
    r = f(rho)
    for k in range(steps):
        s = np.clip(r * s + sg * n[k], -cl, cl).astype(f)     # the device fuses the multiply-add: equal to ~1 ulp
        U[k] = s
    return U.astype(dtype), s.astype(dtype)
