/* brov_oracle.c — CPU ORACLE in plain C (test infrastructure, NOT product code).
 *
 * A scalar float64 restatement of the reference hot path, one vehicle at a time, in the reference's own operation
 * structure (6x6 matrix semantics written out, libm sin/cos/pow, the 3rd-order thruster lag advanced by ONE
 * ThrusterLag.step per dynamics() call — i.e. four sequential lag steps per RK4 step, never a closed form).  OpenMP
 * parallelises over vehicles / windows only.  It exists (a) as a second, independent checker next to
 * oracle/fossen_np.py that is fast enough to verify every vehicle of a 65,536-vehicle GPU rollout, and (b) as the
 * strongest honest CPU baseline (compiled code on all host cores) beside the reference-structured numpy port.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may load this library.  Pinned against
 * tests/golden/reference_vectors.npz (outputs of the unmodified reference) by tests/test_oracle_golden.py.
 *
 * Reference lines restated (ViktorNfa/bluerov2_dynamics):
 *   fossen/BlueROV2.py:23-62 (rotation, Euler-rate matrix with cos(theta) clamp), :245-278 (T200 polynomial with
 *   pow(), lag step, allocation via r x f), :280-355 (C, D, g), :357-400 (dynamics), :503-510 (ThrusterLag.step);
 *   fossen/BlueROV2_wrench.py:27-80,322-367 (quaternion model); training/train_tank_brov2_rk4.py:385-394 (RK4),
 *   training/train_tank_brov2_full_comparison.py:462-465 (Euler), training/train_tank_brov2_wrench_quat.py:262-263.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double m, W, B, xb, yb, zb, Ix, Iy, Iz;
    double added[6], lin[6], quad[6], minv[6], cur[3];
} phys_t;

/* layout of include/brov.h BROV_PH_* (37 doubles) */
static void unpack(const double* ph, phys_t* p) {
    p->m = ph[0]; p->W = ph[1]; p->B = ph[2];
    p->xb = ph[3]; p->yb = ph[4]; p->zb = ph[5];
    p->Ix = ph[6]; p->Iy = ph[7]; p->Iz = ph[8];
    memcpy(p->added, ph + 9, 48); memcpy(p->lin, ph + 15, 48); memcpy(p->quad, ph + 21, 48);
    memcpy(p->minv, ph + 27, 48); memcpy(p->cur, ph + 33, 24);
}

static double sgn(double v) { return (v > 0) - (v < 0); }

static void nu_dot(const phys_t* P, const double* nu, const double R[3][3], const double* tau, double sth,
                   double cs, double cc, double* out) {
    const double u = nu[0], v = nu[1], w = nu[2], p = nu[3], q = nu[4], r = nu[5];
    const double m = P->m;
    const double Xu = P->added[0], Yv = P->added[1], Zw = P->added[2], Kp = P->added[3], Mq = P->added[4], Nr = P->added[5];
    /* C = CRB + CA, entry by entry as fossen/BlueROV2.py:283-325, then C @ nu */
    double C[6][6];
    memset(C, 0, sizeof(C));
    C[0][4] = m * w + -Zw * w;   C[0][5] = -m * v + Yv * v;
    C[1][3] = -m * w + Zw * w;   C[1][5] = m * u + -Xu * u;
    C[2][3] = m * v + -Yv * v;   C[2][4] = -m * u + Xu * u;
    C[3][1] = m * w + -Zw * w;   C[3][2] = -m * v + Yv * v;  C[3][4] = P->Iz * r + -Nr * r;  C[3][5] = -P->Iy * q + Mq * q;
    C[4][0] = -m * w + Zw * w;   C[4][2] = m * u + -Xu * u;  C[4][3] = -P->Iz * r + Nr * r; C[4][5] = P->Ix * p + -Kp * p;
    C[5][0] = m * v + -Yv * v;   C[5][1] = -m * u + Xu * u;  C[5][3] = P->Iy * q + -Mq * q; C[5][4] = -P->Ix * p + Kp * p;
    double nur[6];
    memcpy(nur, nu, 48);
    for (int i = 0; i < 3; ++i) nur[i] -= R[0][i] * P->cur[0] + R[1][i] * P->cur[1] + R[2][i] * P->cur[2];
    const double wmb = P->W - P->B, xbB = P->xb * P->B, ybB = P->yb * P->B, zbB = P->zb * P->B;
    double g[6] = {wmb * sth, -wmb * cs, -wmb * cc, ybB * cc - zbB * cs, -zbB * sth - xbB * cc, xbB * cs + ybB * sth};
    for (int i = 0; i < 6; ++i) {
        double cn = 0;
        for (int j = 0; j < 6; ++j) cn += C[i][j] * nu[j];
        double d = -P->lin[i] - P->quad[i] * fabs(nur[i]);
        out[i] = P->minv[i] * (tau[i] - cn - d * nur[i] - g[i]);
    }
}

static void rhs_euler12(const phys_t* P, const double* x, const double* tau, double* xd) {
    const double cphi = cos(x[3]), sphi = sin(x[3]), cth = cos(x[4]), sth = sin(x[4]), cpsi = cos(x[5]), spsi = sin(x[5]);
    double R[3][3] = {{cpsi * cth, -spsi * cphi + cpsi * sth * sphi, spsi * sphi + cpsi * cphi * sth},
                      {spsi * cth, cpsi * cphi + sphi * sth * spsi, -cpsi * sphi + sth * spsi * cphi},
                      {-sth, cth * sphi, cth * cphi}};
    const double* nu = x + 6;
    for (int i = 0; i < 3; ++i) xd[i] = R[i][0] * nu[0] + R[i][1] * nu[1] + R[i][2] * nu[2];
    double ct = cth;
    if (fabs(ct) < 1e-7) ct = 1e-7 * sgn(ct);
    const double tth = sth / ct;
    xd[3] = nu[3] + sphi * tth * nu[4] + cphi * tth * nu[5];
    xd[4] = cphi * nu[4] - sphi * nu[5];
    xd[5] = sphi / ct * nu[4] + cphi / ct * nu[5];
    nu_dot(P, nu, R, tau, sth, cth * sphi, cth * cphi, xd + 6);
}

static void quat_norm(double* q) {
    double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    if (n < 1e-12) { q[0] = 1; q[1] = q[2] = q[3] = 0; }
    else { for (int i = 0; i < 4; ++i) q[i] /= n; }
}

static void rhs_quat13(const phys_t* P, const double* x, const double* tau, double* xd) {
    double q[4] = {x[3], x[4], x[5], x[6]};
    quat_norm(q);
    const double qw = q[0], qx = q[1], qy = q[2], qz = q[3];
    double R[3][3] = {{1 - 2 * (qy * qy + qz * qz), 2 * (qx * qy - qz * qw), 2 * (qx * qz + qy * qw)},
                      {2 * (qx * qy + qz * qw), 1 - 2 * (qx * qx + qz * qz), 2 * (qy * qz - qx * qw)},
                      {2 * (qx * qz - qy * qw), 2 * (qy * qz + qx * qw), 1 - 2 * (qx * qx + qy * qy)}};
    const double* nu = x + 7;
    for (int i = 0; i < 3; ++i) xd[i] = R[i][0] * nu[0] + R[i][1] * nu[1] + R[i][2] * nu[2];
    const double p = nu[3], qq = nu[4], r = nu[5];
    xd[3] = 0.5 * (-qx * p - qy * qq - qz * r);
    xd[4] = 0.5 * (qw * p + qy * r - qz * qq);
    xd[5] = 0.5 * (qw * qq - qx * r + qz * p);
    xd[6] = 0.5 * (qw * r + qx * qq - qy * p);
    nu_dot(P, nu, R, tau, -R[2][0], R[2][1], R[2][2], xd + 7);
}

typedef struct {
    int model;          /* 0 thruster8, 1 wrench12, 2 quat13 */
    phys_t P;
    double Ad[9], Bd[3], Cc[3];
    double r[8][3], e[8][3];
} model_t;

/* compute_thruster_forces: polynomial (pow, as the reference) -> one lag step per thruster -> sum of [f; r x f] */
static void thruster_forces(const model_t* M, const double* u, double* lag, double* tau) {
    memset(tau, 0, 48);
    for (int i = 0; i < 8; ++i) {
        const double V = u[i];
        const double F = -140.3 * pow(V, 9) + 389.9 * pow(V, 7) - 404.1 * pow(V, 5) + 176.0 * pow(V, 3) + 8.9 * V;
        double* s = lag + 3 * i;
        double n0 = M->Ad[0] * s[0] + M->Ad[1] * s[1] + M->Ad[2] * s[2] + M->Bd[0] * F;
        double n1 = M->Ad[3] * s[0] + M->Ad[4] * s[1] + M->Ad[5] * s[2] + M->Bd[1] * F;
        double n2 = M->Ad[6] * s[0] + M->Ad[7] * s[1] + M->Ad[8] * s[2] + M->Bd[2] * F;
        s[0] = n0; s[1] = n1; s[2] = n2;
        const double Fd = M->Cc[0] * s[0] + M->Cc[1] * s[1] + M->Cc[2] * s[2];
        const double f[3] = {Fd * M->e[i][0], Fd * M->e[i][1], Fd * M->e[i][2]};
        const double* rr = M->r[i];
        tau[0] += f[0]; tau[1] += f[1]; tau[2] += f[2];
        tau[3] += rr[1] * f[2] - rr[2] * f[1];
        tau[4] += rr[2] * f[0] - rr[0] * f[2];
        tau[5] += rr[0] * f[1] - rr[1] * f[0];
    }
}

static void dynamics(const model_t* M, const double* x, const double* u, double* lag, double* xd) {
    if (M->model == 0) {
        double tau[6];
        thruster_forces(M, u, lag, tau);
        rhs_euler12(&M->P, x, tau, xd);
    } else if (M->model == 1) {
        rhs_euler12(&M->P, x, u, xd);
    } else {
        rhs_quat13(&M->P, x, u, xd);
    }
}

static void step(const model_t* M, int euler, double dt, double* x, const double* u, double* lag) {
    const int nx = M->model == 2 ? 13 : 12;
    double k1[13], k2[13], k3[13], k4[13], t[13];
    if (euler) {
        dynamics(M, x, u, lag, k1);
        for (int i = 0; i < nx; ++i) x[i] = x[i] + dt * k1[i];
    } else {
        dynamics(M, x, u, lag, k1);
        for (int i = 0; i < nx; ++i) t[i] = x[i] + 0.5 * dt * k1[i];
        dynamics(M, t, u, lag, k2);
        for (int i = 0; i < nx; ++i) t[i] = x[i] + 0.5 * dt * k2[i];
        dynamics(M, t, u, lag, k3);
        for (int i = 0; i < nx; ++i) t[i] = x[i] + dt * k3[i];
        dynamics(M, t, u, lag, k4);
        for (int i = 0; i < nx; ++i) x[i] = x[i] + (dt / 6.0) * (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]);
    }
    if (M->model == 2) quat_norm(x + 3);
}

static void make_model(model_t* M, int model, const double* phys, const double* Ad, const double* Bd, const double* r,
                       const double* e) {
    M->model = model;
    unpack(phys, &M->P);
    if (Ad) memcpy(M->Ad, Ad, 72);
    if (Bd) memcpy(M->Bd, Bd, 24);
    M->Cc[0] = 0.0; M->Cc[1] = 5.992; M->Cc[2] = 3.317;
    if (r) memcpy(M->r, r, sizeof(M->r));
    if (e) memcpy(M->e, e, sizeof(M->e));
}

/* Rollout of n vehicles.  phys: [37] shared or [n][37] per vehicle (per_vehicle != 0).  U: time-major [T][n][nu]
 * (u_shared = 0) or [T][nu].  lag: [n][24] in/out (thruster model; may be NULL = zero, discarded).
 * traj: [T/stride][n][nx] or NULL.  Ad, Bd, r, e: lag discretisation and thruster geometry (thruster model).
 * min_abs_cos: [n] in/out or NULL — running minimum of |cos theta| over the states the steps START from (Euler-angle
 * models): the distance to the singularity of euler_kinematics_matrix (fossen/BlueROV2.py:43-62). */
int brov_oracle_rollout(int model, int euler, long long n, long long T, double dt, const double* phys, int per_vehicle,
                        const double* Ad, const double* Bd, const double* r, const double* e, double* x,
                        const double* U, int u_shared, double* lag, double* traj, long long stride,
                        double* min_abs_cos) {
    const int nx = model == 2 ? 13 : 12, nu = model == 0 ? 8 : 6;
#pragma omp parallel for schedule(static)
    for (long long i = 0; i < n; ++i) {
        model_t M;
        make_model(&M, model, phys + (per_vehicle ? 37 * i : 0), Ad, Bd, r, e);
        double xs[13], ls[24];
        memcpy(xs, x + i * nx, sizeof(double) * nx);
        if (lag) memcpy(ls, lag + 24 * i, sizeof(ls)); else memset(ls, 0, sizeof(ls));
        double mc = min_abs_cos ? min_abs_cos[i] : 1.0;
        for (long long k = 0; k < T; ++k) {
            const double* u = u_shared ? U + k * nu : U + (k * n + i) * nu;
            if (min_abs_cos && model != 2) { const double c = fabs(cos(xs[4])); if (c < mc) mc = c; }
            step(&M, euler, dt, xs, u, ls);
            if (traj && stride > 0 && (k + 1) % stride == 0)
                memcpy(traj + (((k + 1) / stride - 1) * n + i) * nx, xs, sizeof(double) * nx);
        }
        memcpy(x + i * nx, xs, sizeof(double) * nx);
        if (lag) memcpy(lag + 24 * i, ls, sizeof(ls));
        if (min_abs_cos) min_abs_cos[i] = mc;
    }
    return 0;
}

/* multistep_rmse_endpoint_physics with every window starting from zero lag ("reset"): returns sum of squared endpoint
 * errors over windows k = 0..rows-H-1. */
double brov_oracle_multistep_se(int model, int euler, long long rows, long long H, double dt, const double* phys,
                                const double* Ad, const double* Bd, const double* r, const double* e, const double* X,
                                const double* U) {
    const int nx = model == 2 ? 13 : 12, nu = model == 0 ? 8 : 6;
    const long long ns = rows - H;
    double se = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : se)
    for (long long k = 0; k < ns; ++k) {
        model_t M;
        make_model(&M, model, phys, Ad, Bd, r, e);
        double xs[13], ls[24];
        memcpy(xs, X + k * nx, sizeof(double) * nx);
        memset(ls, 0, sizeof(ls));
        for (long long j = 0; j < H; ++j) step(&M, euler, dt, xs, U + (k + j) * nu, ls);
        for (int i = 0; i < nx; ++i) { double d = xs[i] - X[(k + H) * nx + i]; se += d * d; }
    }
    return se;
}

int brov_oracle_threads(void) {
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    return omp_get_max_threads();
#else
    return 1;
#endif
}
