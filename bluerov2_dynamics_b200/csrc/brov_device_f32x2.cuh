// brov_device_f32x2.cuh — packed-FP32 RK4 step of the 8-thruster model (sm_100 FFMA2 / FMUL2 / FADD2).
//
// The fp32 rollout is ISSUE-bound: 972 warp instructions per RK4 step, 787 of them FP32, at 88 % of the issue slots
// (profiles/r01b_opmix.txt).  Blackwell's packed instructions perform two independent FP32 operations on a 64-bit
// register pair in one issue slot (same FP32-pipe cycles), so everything that is "the same operation on six axes / eight
// thrusters / twelve states" is done on pairs here: the T200 polynomials (4 pairs), the allocation (pair products +
// one horizontal add per row), the closed-form lag filter (3 axis pairs x 3 states), damping and the final assembly of
// nu_dot, the RK4 accumulation and stage states (6 state pairs) and the angle-addition trig of (theta, psi).  The
// irregular parts (rotation R nu, Euler rates, Coriolis cross terms, restoring terms) stay scalar on the halves of the
// same registers — unpacking is free.  Same arithmetic as the scalar path up to the order of a few additions.
//
// Used for <float, THRUSTER8, RK4, projected lag, constants shared by all vehicles>; every other instantiation keeps
// the scalar step.  Reference lines restated: see brov_device.cuh.
#pragma once

namespace brov {

struct f2 { unsigned long long v; };
__device__ __forceinline__ f2 mk2(float a, float b) { f2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void un2(f2 p, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(p.v)); }
__device__ __forceinline__ float lo2(f2 p) { float a, b; un2(p, a, b); return a; }
__device__ __forceinline__ float hi2(f2 p) { float a, b; un2(p, a, b); return b; }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { f2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { f2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { f2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
__device__ __forceinline__ f2 sub2(f2 a, f2 b) { f2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }

// layout of Consts<float>::pk in PAIRS (filled by make_consts in brov_api.cu)
enum : int {
    PK_POLY = 0,     // 5 splats: -140.3, 389.9, -404.1, 176.0, 8.9
    PK_LAGG = 5,     // per stage j (4): splat G_j0, G_j1, G_j2, H_j          -> 16 pairs
    PK_LAGA = 21,    // per state row k (3): splat A_k0, A_k1, A_k2, B_k       -> 12 pairs
    PK_DL = 33,      // 3 axis pairs of linear damping
    PK_DQ = 36,      // 3 axis pairs of quadratic damping
    PK_MINV = 39,    // 3 axis pairs of 1/M
    PK_ROT = 42,     // splats: sin  1/120, -1/6 ; cos -1/720, 1/24, -1/2, 1   -> 6 pairs
    PK_RK = 48,      // splats: dt/2, dt, dt/6, 2                              -> 4 pairs
    PK_PAIRS = 52
};

// angle-addition sin/cos of two angles at once: (s, c) of base + d from the base values (see rotate_sc)
__device__ __forceinline__ void rotate_sc2(const f2* __restrict__ K, f2 s, f2 c, f2 ns, f2 d, f2* so, f2* co) {
    const f2 z = mul2(d, d);
    const f2 sd = fma2(mul2(d, z), fma2(z, K[PK_ROT + 0], K[PK_ROT + 1]), d);
    const f2 cd = fma2(z, fma2(z, fma2(z, K[PK_ROT + 2], K[PK_ROT + 3]), K[PK_ROT + 4]), K[PK_ROT + 5]);
    *so = fma2(s, cd, mul2(c, sd));
    *co = fma2(c, cd, mul2(ns, sd));
}

// state derivative on pairs.  xp: state pairs (x,y) (z,phi) (theta,psi) (u,v) (w,p) (q,r); tr: sin/cos of the angles;
// taup: body wrench pairs; kd: derivative pairs out.
__device__ __forceinline__ void rhs_packed(const Consts<float>& c, const f2* __restrict__ K, const f2* __restrict__ xp,
                                           const Trig<float>& tr, const f2* __restrict__ taup, f2* __restrict__ kd) {
    const float* p = c.kp;
    float u, v, w, pp, q, r;
    un2(xp[3], u, v);
    un2(xp[4], w, pp);
    un2(xp[5], q, r);
    const float sphi = tr.sphi, cphi = tr.cphi, sth = tr.sth, cth = tr.cth, spsi = tr.spsi, cpsi = tr.cpsi;
    // p_dot = Rz Ry Rx nu_1 as three planar rotations
    const float v1 = cphi * v - sphi * w;
    const float w1 = sphi * v + cphi * w;
    const float u2 = cth * u + sth * w1;
    const float zd = cth * w1 - sth * u;
    kd[0] = mk2(cpsi * u2 - spsi * v1, spsi * u2 + cpsi * v1);
    // Euler rates with the reference's cos(theta) clamp
    float ct = cth;
    if (fabsf(ct) < 1e-7f) ct = (ct > 0.0f) ? 1e-7f : ((ct < 0.0f) ? -1e-7f : 0.0f);
    const float ic = rcp_(ct);
    const float psd = (sphi * q + cphi * r) * ic;
    kd[1] = mk2(zd, pp + sth * psd);
    kd[2] = mk2(cphi * q - sphi * r, psd);
    // relative velocity
    f2 nr0 = xp[3], nr1 = xp[4];
    const f2 nr2 = xp[5];
    if (c.has_current) {
        const float cx = p[KP_CUR + 0], cy = p[KP_CUR + 1], cz = p[KP_CUR + 2];
        const float x1 = cpsi * cx + spsi * cy;
        const float y1 = cpsi * cy - spsi * cx;
        const float x2 = cth * x1 - sth * cz;
        const float z2 = sth * x1 + cth * cz;
        nr0 = mk2(u - x2, v - (cphi * y1 + sphi * z2));
        nr1 = mk2(w - (cphi * z2 - sphi * y1), pp);
    }
    // Coriolis (closed form) + restoring, scalar: cg = C(nu) nu + g(eta)
    const float cs = cth * sphi, cc = cth * cphi;
    const float a1u = p[KP_A + 0] * u, a2v = p[KP_A + 1] * v, a3w = p[KP_A + 2] * w;
    const float wmb = p[KP_WMB], xbB = p[KP_XBB + 0], ybB = p[KP_XBB + 1], zbB = p[KP_XBB + 2];
    const float cg0 = fmaf(wmb, sth, a3w * q - a2v * r);
    const float cg1 = fmaf(-wmb, cs, a1u * r - a3w * pp);
    const float cg2 = fmaf(-wmb, cc, a2v * pp - a1u * q);
    const float cg3 = fmaf(ybB, cc, fmaf(-zbB, cs, p[KP_DA + 0] * (v * w) + p[KP_DB + 0] * (q * r)));
    const float cg4 = fmaf(-zbB, sth, fmaf(-xbB, cc, p[KP_DA + 1] * (u * w) + p[KP_DB + 1] * (pp * r)));
    const float cg5 = fmaf(xbB, cs, fmaf(ybB, sth, p[KP_DA + 2] * (u * v) + p[KP_DB + 2] * (pp * q)));
    // damping and assembly on axis pairs: nu_dot = Minv (tau - cg - (DL + DQ |nu_r|) nu_r)
    float a, b;
    un2(nr0, a, b);
    const f2 t0 = mk2(a * fabsf(a), b * fabsf(b));
    un2(nr1, a, b);
    const f2 t1 = mk2(a * fabsf(a), b * fabsf(b));
    const f2 t2 = mk2(q * fabsf(q), r * fabsf(r));
    const f2 e0 = fma2(K[PK_DQ + 0], t0, mul2(K[PK_DL + 0], nr0));
    const f2 e1 = fma2(K[PK_DQ + 1], t1, mul2(K[PK_DL + 1], nr1));
    const f2 e2 = fma2(K[PK_DQ + 2], t2, mul2(K[PK_DL + 2], nr2));
    kd[3] = mul2(K[PK_MINV + 0], sub2(sub2(taup[0], mk2(cg0, cg1)), e0));
    kd[4] = mul2(K[PK_MINV + 1], sub2(sub2(taup[1], mk2(cg2, cg3)), e1));
    kd[5] = mul2(K[PK_MINV + 2], sub2(sub2(taup[2], mk2(cg4, cg5)), e2));
}

// One RK4 step.  x[12] state, Z[18] allocation-projected lag state (index 3 r + k), u[8] thruster voltages.
__device__ __forceinline__ void integrate_step_packed(const Consts<float>& c, float* __restrict__ x,
                                                      float* __restrict__ Z, const float* __restrict__ u) {
    const f2* K = reinterpret_cast<const f2*>(c.pk);
    const f2* AL = reinterpret_cast<const f2*>(&c.alloc[0][0]);   // AL[4 r + i / 2] = (alloc[r][i], alloc[r][i + 1])
    // static thrust of the 8 thrusters, Horner in V^2 on 4 pairs
    f2 F[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const f2 V = mk2(u[2 * i], u[2 * i + 1]);
        const f2 z = mul2(V, V);
        f2 q = fma2(K[PK_POLY + 0], z, K[PK_POLY + 1]);
        q = fma2(q, z, K[PK_POLY + 2]);
        q = fma2(q, z, K[PK_POLY + 3]);
        q = fma2(q, z, K[PK_POLY + 4]);
        F[i] = mul2(q, V);
    }
    // allocation tau_F = alloc F: pair products, one horizontal add per row (structural zeros skipped)
    float tf[6];
    {
        float a, b;
        un2(fma2(AL[0 * 4 + 1], F[1], mul2(AL[0 * 4 + 0], F[0])), a, b); tf[0] = a + b;
        un2(fma2(AL[1 * 4 + 1], F[1], mul2(AL[1 * 4 + 0], F[0])), a, b); tf[1] = a + b;
        un2(fma2(AL[2 * 4 + 3], F[3], mul2(AL[2 * 4 + 2], F[2])), a, b); tf[2] = a + b;
        un2(fma2(AL[3 * 4 + 3], F[3], fma2(AL[3 * 4 + 2], F[2], fma2(AL[3 * 4 + 1], F[1], mul2(AL[3 * 4 + 0], F[0])))), a, b); tf[3] = a + b;
        un2(fma2(AL[4 * 4 + 3], F[3], fma2(AL[4 * 4 + 2], F[2], fma2(AL[4 * 4 + 1], F[1], mul2(AL[4 * 4 + 0], F[0])))), a, b); tf[4] = a + b;
        un2(fma2(AL[5 * 4 + 1], F[1], mul2(AL[5 * 4 + 0], F[0])), a, b); tf[5] = a + b;
    }
    const f2 Fu[3] = {mk2(tf[0], tf[1]), mk2(tf[2], tf[3]), mk2(tf[4], tf[5])};
    // lag state as axis pairs: Zp[k][rp] = (Z[2 rp][k], Z[2 rp + 1][k])
    f2 Zp[3][3];
#pragma unroll
    for (int rp = 0; rp < 3; ++rp)
#pragma unroll
        for (int k = 0; k < 3; ++k) Zp[k][rp] = mk2(Z[6 * rp + k], Z[6 * rp + 3 + k]);
    // wrench seen by the four dynamics() calls of the step (closed form of the lag over its sub-steps)
    auto wrench = [&](int j, f2* tau) {
#pragma unroll
        for (int rp = 0; rp < 3; ++rp)
            tau[rp] = fma2(K[PK_LAGG + 4 * j + 0], Zp[0][rp],
                           fma2(K[PK_LAGG + 4 * j + 1], Zp[1][rp],
                                fma2(K[PK_LAGG + 4 * j + 2], Zp[2][rp], mul2(K[PK_LAGG + 4 * j + 3], Fu[rp]))));
    };

    f2 xp[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) xp[i] = mk2(x[2 * i], x[2 * i + 1]);
    Trig<float> tr0;
    trig_full<float>(x + 3, tr0);
    const f2 sb = mk2(tr0.sth, tr0.spsi), cb = mk2(tr0.cth, tr0.cpsi), nsb = mk2(-tr0.sth, -tr0.spsi);

    f2 k[6], tau[3], acc[6], xs[6];
    wrench(0, tau);
    rhs_packed(c, K, xp, tr0, tau, k);
#pragma unroll
    for (int s = 1; s <= 3; ++s) {
        const f2 h = (s == 3) ? K[PK_RK + 1] : K[PK_RK + 0];
#pragma unroll
        for (int i = 0; i < 6; ++i) {
            if (s == 1) acc[i] = k[i]; else acc[i] = fma2(K[PK_RK + 3], k[i], acc[i]);
            xs[i] = fma2(h, k[i], xp[i]);
        }
        // stage angles = base + d: phi scalar, (theta, psi) as a pair
        const f2 d23 = mul2(h, k[2]);
        const float dphi = lo2(h) * hi2(k[1]);
        float dth, dpsi;
        un2(d23, dth, dpsi);
        Trig<float> trs;
        if (fmaxf(fmaxf(fabsf(dphi), fabsf(dth)), fabsf(dpsi)) > 0.125f) {
            const float ang[3] = {hi2(xs[1]), lo2(xs[2]), hi2(xs[2])};
            trig_full<float>(ang, trs);
        } else {
            rotate_sc(tr0.sphi, tr0.cphi, dphi, &trs.sphi, &trs.cphi);
            f2 so, co;
            rotate_sc2(K, sb, cb, nsb, d23, &so, &co);
            un2(so, trs.sth, trs.spsi);
            un2(co, trs.cth, trs.cpsi);
        }
        wrench(s, tau);
        rhs_packed(c, K, xs, trs, tau, k);
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        xp[i] = fma2(K[PK_RK + 2], add2(acc[i], k[i]), xp[i]);
        un2(xp[i], x[2 * i], x[2 * i + 1]);
    }
    // lag state after the four sub-steps
#pragma unroll
    for (int rp = 0; rp < 3; ++rp) {
        const f2 a = Zp[0][rp], b = Zp[1][rp], d = Zp[2][rp];
#pragma unroll
        for (int kk = 0; kk < 3; ++kk) {
            const f2 nz = fma2(K[PK_LAGA + 4 * kk + 0], a,
                               fma2(K[PK_LAGA + 4 * kk + 1], b,
                                    fma2(K[PK_LAGA + 4 * kk + 2], d, mul2(K[PK_LAGA + 4 * kk + 3], Fu[rp]))));
            un2(nz, Z[6 * rp + kk], Z[6 * rp + 3 + kk]);
        }
    }
}

}  // namespace brov
