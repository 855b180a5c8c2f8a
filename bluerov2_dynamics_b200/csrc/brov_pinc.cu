// brov_pinc.cu — inference rollout of the reference's PINc residual network on sm_100a.
//
// Reference (ViktorNfa/bluerov2_dynamics, training/train_tank_brov2_rk4.py):
//   thrusters_to_body_wrenches  :553-562   u8 -> compute_thruster_forces (polynomial + ONE lag step) -> [X, Y, Z, Mz]
//   dataset12_to_9 / state9_to_12 :565-596 [x y z cos(psi) sin(psi) u v w r] <-> 12-state rows
//   AdaptiveSoftplus            :596-604   softplus(beta x) / (beta + 1e-12)
//   PINcNet                     :607-673   14 -> 64 -> 64 -> 64 -> 64 -> 9, (Linear, AdaptiveSoftplus, LayerNorm) x 4 +
//                                          Linear; x_next = x + dx with dx_xy rotated by the current yaw and
//                                          (cos, sin) re-normalised
//   simulate_pinc               :789-812   per step: thruster map (float64) -> float32 network
//   multistep_rmse_endpoint_pinc :815-840  sliding windows, endpoint error in the 12-state projection
//
// One thread per window.  The 55 KB of network weights sit in shared memory once per block and are read as
// broadcast 128-bit loads; a thread keeps the 64 pre-activations of the layer being built in registers and its
// activations in a private shared-memory column ([unit][thread], conflict-free).  The thruster map runs inline in
// float64 on the four allocation-projected lag filters the network input needs (rows X, Y, Z, Mz of the allocation
// matrix: 12 hidden values instead of the reference's 24), exactly one lag step per rollout step like the reference.
#include <cuda_fp16.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "brov_internal.cuh"

namespace {

constexpr int HID = 64;
constexpr int NIN = 14;
constexpr int PB = 384;   // threads per block: 12 warps share one 58 KB copy of the weights; with the 96 KB of activation
                          // columns that is one block per SM, and 168 registers x 384 threads fill the register file

// packed weight blob (floats), see brov_pinc_create
constexpr int OFF_W0 = 0;                         // [14][64]  (input-major: the 64 weights of one input contiguous)
constexpr int OFF_B0 = OFF_W0 + NIN * HID;        // b, ln_w, ln_b of layer 0: 3 x [64]
constexpr int OFF_L1 = OFF_B0 + 3 * HID;          // layers 1..3: [64][64] + 3 x [64] each
constexpr int LSTRIDE = HID * HID + 3 * HID;
constexpr int OFF_W4 = OFF_L1 + 3 * LSTRIDE;      // [64][12]  (hidden-major, 9 outputs padded to 12)
constexpr int OFF_B4 = OFF_W4 + HID * 12;         // [12]
constexpr int NW = OFF_B4 + 12;

struct ThrMap {   // single-step thruster map constants (float64)
    double Ad[9], Bd[3], Cc[3];
    double a4[4][8];   // rows X, Y, Z, Mz of the allocation matrix
    double dt;
};

struct PincParams {
    const float* w;     // packed blob, device
    float beta[4];
};

// Packed FP32 arithmetic of sm_100: one FFMA2 instruction performs two fused multiply-adds on a 64-bit register pair,
// halving the issue slots of the dense layers (the FP32 pipe rate is unchanged; the freed slots go to the shared-memory
// loads that feed it).
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack2(float a, float b) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& a, float& b) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v));
}
__device__ __forceinline__ void fma2(u64& d, u64 a, u64 b) {   // d += a * b, lane-wise on (lo, hi)
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
}

// torch.nn.functional.softplus(beta = 1, threshold = 20) = x > 20 ? x : log1p(exp(x)), evaluated as
// max(x, 0) + log(1 + exp(-|x|)) with the two MUFU approximations: absolute error <= 1.2e-7 (the rounding of 1 + e),
// i.e. float32 resolution of the O(1) activations that LayerNorm is about to standardise.
__device__ __forceinline__ float softplus_f(float x) {
    float e, l;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(fabsf(x) * -1.4426950408889634f));   // exp(-|x|)
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));                           // argument in [1, 2]
    return fmaf(l, 0.6931471805599453f, fmaxf(x, 0.0f));
}

// activation + LayerNorm of the 64 pre-activations (32 packed pairs), result to the thread's shared-memory column
// (element stride PBLK)
template <int PBLK>
__device__ __forceinline__ void act_norm_store(const u64* a2, float beta, const float* __restrict__ lnw,
                                               const float* __restrict__ lnb, float* __restrict__ hcol) {
    const float ib = 1.0f / (beta + 1e-12f);
    float a[HID];
    float mean = 0.0f;
#pragma unroll
    for (int j = 0; j < HID / 2; ++j) {
        float lo, hi;
        unpack2(a2[j], lo, hi);
        a[2 * j] = softplus_f(beta * lo) * ib;
        a[2 * j + 1] = softplus_f(beta * hi) * ib;
        mean += a[2 * j] + a[2 * j + 1];
    }
    mean *= (1.0f / HID);
    float var = 0.0f;
#pragma unroll
    for (int j = 0; j < HID; ++j) {
        a[j] -= mean;
        var = fmaf(a[j], a[j], var);
    }
    const float rstd = rsqrtf(var * (1.0f / HID) + 1e-5f);
#pragma unroll
    for (int q = 0; q < HID / 4; ++q) {
        const float4 g = *reinterpret_cast<const float4*>(lnw + 4 * q);
        const float4 b = *reinterpret_cast<const float4*>(lnb + 4 * q);
        hcol[(4 * q + 0) * PBLK] = fmaf(a[4 * q + 0] * rstd, g.x, b.x);
        hcol[(4 * q + 1) * PBLK] = fmaf(a[4 * q + 1] * rstd, g.y, b.y);
        hcol[(4 * q + 2) * PBLK] = fmaf(a[4 * q + 2] * rstd, g.z, b.z);
        hcol[(4 * q + 3) * PBLK] = fmaf(a[4 * q + 3] * rstd, g.w, b.w);
    }
}

template <int NPAIR, int NWIN>
__device__ __forceinline__ void bias_init(u64 (&a2)[NWIN][NPAIR], const float* __restrict__ b) {
#pragma unroll
    for (int q = 0; q < NPAIR / 2; ++q) {
        const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(b + 4 * q);
#pragma unroll
        for (int w = 0; w < NWIN; ++w) {
            a2[w][2 * q] = v.x;
            a2[w][2 * q + 1] = v.y;
        }
    }
}

// a2[w] += wrow * x[w] for NPAIR packed pairs and NWIN windows: every 128-bit broadcast load of four weights from
// shared memory feeds 2 NWIN packed FMAs (the shared-memory pipe is the bottleneck at one window per thread:
// 90 % of its peak in profiles/r01h_compare_raw.csv)
template <int NPAIR, int NWIN>
__device__ __forceinline__ void axpy(u64 (&a2)[NWIN][NPAIR], const float* __restrict__ wrow, const float (&x)[NWIN]) {
    u64 xx[NWIN];
#pragma unroll
    for (int w = 0; w < NWIN; ++w) xx[w] = pack2(x[w], x[w]);
#pragma unroll
    for (int q = 0; q < NPAIR / 2; ++q) {
        const ulonglong2 wv = *reinterpret_cast<const ulonglong2*>(wrow + 4 * q);
#pragma unroll
        for (int w = 0; w < NWIN; ++w) {
            fma2(a2[w][2 * q], wv.x, xx[w]);
            fma2(a2[w][2 * q + 1], wv.y, xx[w]);
        }
    }
}

// PINcNet.forward for NWIN windows of one thread: z[w][14] -> xn[w][9].  sw: weights in shared memory; hbase: this
// thread's activation columns, element (window w, unit j) at hbase[(w * HID + j) * PBLK].
template <int NWIN, int PBLK>
__device__ __forceinline__ void pinc_forward(const float* __restrict__ sw, const float* beta, float* __restrict__ hbase,
                                             const float (&z)[NWIN][NIN], float (&xn)[NWIN][9]) {
    u64 a2[NWIN][HID / 2];
    float xin[NWIN];
    bias_init<HID / 2, NWIN>(a2, sw + OFF_B0);
#pragma unroll
    for (int i = 0; i < NIN; ++i) {
#pragma unroll
        for (int w = 0; w < NWIN; ++w) xin[w] = z[w][i];
        axpy<HID / 2, NWIN>(a2, sw + OFF_W0 + i * HID, xin);
    }
#pragma unroll
    for (int w = 0; w < NWIN; ++w)
        act_norm_store<PBLK>(a2[w], beta[0], sw + OFF_B0 + HID, sw + OFF_B0 + 2 * HID, hbase + w * HID * PBLK);
#pragma unroll 1
    for (int l = 0; l < 3; ++l) {
        const float* L = sw + OFF_L1 + l * LSTRIDE;
        bias_init<HID / 2, NWIN>(a2, L + HID * HID);
#pragma unroll 4
        for (int i = 0; i < HID; ++i) {
#pragma unroll
            for (int w = 0; w < NWIN; ++w) xin[w] = hbase[(w * HID + i) * PBLK];
            axpy<HID / 2, NWIN>(a2, L + i * HID, xin);
        }
#pragma unroll
        for (int w = 0; w < NWIN; ++w)
            act_norm_store<PBLK>(a2[w], beta[l + 1], L + HID * HID + HID, L + HID * HID + 2 * HID, hbase + w * HID * PBLK);
    }
    u64 d2[NWIN][6];
    bias_init<6, NWIN>(d2, sw + OFF_B4);
#pragma unroll 4
    for (int i = 0; i < HID; ++i) {
#pragma unroll
        for (int w = 0; w < NWIN; ++w) xin[w] = hbase[(w * HID + i) * PBLK];
        axpy<6, NWIN>(d2, sw + OFF_W4 + i * 12, xin);
    }
#pragma unroll
    for (int w = 0; w < NWIN; ++w) {
        float dx[12];
#pragma unroll
        for (int q = 0; q < 6; ++q) unpack2(d2[w][q], dx[2 * q], dx[2 * q + 1]);
        // residual update; body-frame (dx, dy) rotated by the CURRENT yaw; (cos, sin) re-normalised (:639-673)
        const float c = z[w][3], s = z[w][4];
        float base[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) base[j] = z[w][j] + dx[j];
        xn[w][0] = (c * dx[0] - s * dx[1]) + z[w][0];
        xn[w][1] = (s * dx[0] + c * dx[1]) + z[w][1];
        xn[w][2] = base[2];
        const float nrm = fmaxf(sqrtf(base[3] * base[3] + base[4] * base[4]), 1e-6f);
        xn[w][3] = base[3] / nrm;
        xn[w][4] = base[4] / nrm;
#pragma unroll
        for (int j = 5; j < 9; ++j) xn[w][j] = base[j];
    }
}

__device__ __forceinline__ double poly_t200(double V) {
    const double z = V * V;
    double p = fma(-140.3, z, 389.9);
    p = fma(p, z, -404.1);
    p = fma(p, z, 176.0);
    p = fma(p, z, 8.9);
    return p * V;
}

// one thruster-map step on the projected lag Z[4][3]: returns u4 = [X, Y, Z, Mz] AFTER the lag update (the reference's
// ThrusterLag.step returns C x of the updated state)
__device__ __forceinline__ void thruster_map4(const ThrMap& m, const double* __restrict__ u8, double* __restrict__ Z,
                                              double* __restrict__ u4) {
    double F[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) F[i] = poly_t200(u8[i]);
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        double tf = 0.0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool nz = (r < 2 || r == 3) ? (i < 4) : (i >= 4);  // rows X, Y, Mz: horizontal thrusters; Z: vertical
            if (nz) tf = fma(m.a4[r][i], F[i], tf);
        }
        const double a = Z[3 * r], b = Z[3 * r + 1], c = Z[3 * r + 2];
        const double n0 = m.Ad[0] * a + m.Ad[1] * b + m.Ad[2] * c + m.Bd[0] * tf;
        const double n1 = m.Ad[3] * a + m.Ad[4] * b + m.Ad[5] * c + m.Bd[1] * tf;
        const double n2 = m.Ad[6] * a + m.Ad[7] * b + m.Ad[8] * c + m.Bd[2] * tf;
        Z[3 * r] = n0; Z[3 * r + 1] = n1; Z[3 * r + 2] = n2;
        u4[r] = m.Cc[0] * n0 + m.Cc[1] * n1 + m.Cc[2] * n2;
    }
}

// lag state [8][3] of the reference -> projected [4][3]
__device__ __forceinline__ void project4(const ThrMap& m, const double* __restrict__ lag24, double* __restrict__ Z) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < 8; ++i) s = fma(m.a4[r][i], lag24[3 * i + k], s);
            Z[3 * r + k] = s;
        }
}


__device__ __forceinline__ void x12_to_9(const double* __restrict__ x12, float* __restrict__ x9) {
    double s, c;
    sincos(x12[5], &s, &c);
    x9[0] = (float)x12[0]; x9[1] = (float)x12[1]; x9[2] = (float)x12[2];
    x9[3] = (float)c; x9[4] = (float)s;
    x9[5] = (float)x12[6]; x9[6] = (float)x12[7]; x9[7] = (float)x12[8]; x9[8] = (float)x12[11];
}
__device__ __forceinline__ void x9_to_12(const float* __restrict__ x9, double* __restrict__ x12) {
    x12[0] = x9[0]; x12[1] = x9[1]; x12[2] = x9[2];
    x12[3] = 0.0; x12[4] = 0.0; x12[5] = atan2((double)x9[4], (double)x9[3]);
    x12[6] = x9[5]; x12[7] = x9[6]; x12[8] = x9[7];
    x12[9] = 0.0; x12[10] = 0.0; x12[11] = x9[8];
}

// ---------------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------------
constexpr int PB2 = 256;   // threads per block of the rollout / scoring kernels, TWO windows per thread
constexpr int NWIN = 2;
constexpr size_t SMEM_BYTES = (size_t)(NW + HID * PB) * sizeof(float);                 // forward kernel
constexpr size_t SMEM_BYTES2 = (size_t)(NW + NWIN * HID * PB2) * sizeof(float);        // 190 KB: one block per SM

__device__ __forceinline__ void load_weights_n(const float* __restrict__ g, float* __restrict__ sw, int nthreads) {
    for (int e = threadIdx.x * 4; e < NW; e += nthreads * 4)
        *reinterpret_cast<float4*>(sw + e) = *reinterpret_cast<const float4*>(g + e);
}

__global__ void __launch_bounds__(PB, 1) pinc_forward_kernel(PincParams p, const float* __restrict__ Zin,
                                                          float* __restrict__ out, long long n) {
    extern __shared__ __align__(16) float smf[];
    float* sw = smf;
    float* hcol = smf + NW + threadIdx.x;
    load_weights_n(p.w, sw, PB);
    __syncthreads();
    const long long i = (long long)blockIdx.x * PB + threadIdx.x;
    if (i >= n) return;
    float z[1][NIN], xn[1][9];
#pragma unroll
    for (int j = 0; j < NIN; ++j) z[0][j] = Zin[i * NIN + j];
    pinc_forward<1, PB>(sw, p.beta, hcol, z, xn);
#pragma unroll
    for (int j = 0; j < 9; ++j) out[i * 9 + j] = xn[0][j];
}

struct PincRollArgs {
    PincParams p;
    ThrMap m;
    const double* x0;      // [n][12]
    const double* U;       // element (k, i, j) at U[k*u_stride_t + i*u_stride_n + j]
    long long u_stride_t, u_stride_n;
    const double* lag_in;  // [n][24] or nullptr
    double* lag_out;       // [n][12] projected (X, Y, Z, Mz rows) or nullptr
    double* traj;          // [steps/stride][n][12] or nullptr
    float* x9T;            // [n][9] or nullptr
    long long n;
    int steps, stride;
};

// Thread t of block b rolls windows i0 = b * 2 PB2 + t and i1 = i0 + PB2.  A thread whose second (or both) window
// falls off the end shadows the last window and stores nothing for it.
__global__ void __launch_bounds__(PB2, 1) pinc_rollout_kernel(const __grid_constant__ PincRollArgs a) {
    extern __shared__ __align__(16) float smf[];
    float* sw = smf;
    float* hbase = smf + NW + threadIdx.x;
    load_weights_n(a.p.w, sw, PB2);
    __syncthreads();
    long long idx[NWIN];
    bool ok[NWIN];
    float z[NWIN][NIN], xn[NWIN][9];
    double Z[NWIN][12];
#pragma unroll
    for (int w = 0; w < NWIN; ++w) {
        const long long i = (long long)blockIdx.x * (NWIN * PB2) + w * PB2 + threadIdx.x;
        ok[w] = i < a.n;
        idx[w] = ok[w] ? i : a.n - 1;
        double x12[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) x12[j] = a.x0[idx[w] * 12 + j];
        x12_to_9(x12, z[w]);
        z[w][13] = (float)a.m.dt;
        if (a.lag_in) {
            double l24[24];
#pragma unroll
            for (int j = 0; j < 24; ++j) l24[j] = a.lag_in[idx[w] * 24 + j];
            project4(a.m, l24, Z[w]);
        } else {
#pragma unroll
            for (int j = 0; j < 12; ++j) Z[w][j] = 0.0;
        }
    }
    for (int k = 0; k < a.steps; ++k) {
#pragma unroll
        for (int w = 0; w < NWIN; ++w) {
            double u8[8], u4[4];
            const double* up = a.U + idx[w] * a.u_stride_n + (long long)k * a.u_stride_t;
#pragma unroll
            for (int j = 0; j < 8; ++j) u8[j] = __ldg(up + j);
            thruster_map4(a.m, u8, Z[w], u4);
#pragma unroll
            for (int j = 0; j < 4; ++j) z[w][9 + j] = (float)u4[j];
        }
        pinc_forward<NWIN, PB2>(sw, a.p.beta, hbase, z, xn);
        const bool snap = a.traj && (k + 1) % a.stride == 0;
#pragma unroll
        for (int w = 0; w < NWIN; ++w) {
#pragma unroll
            for (int j = 0; j < 9; ++j) z[w][j] = xn[w][j];
            if (snap && ok[w]) {
                double x12[12];
                x9_to_12(xn[w], x12);
                double* dst = a.traj + (((long long)(k + 1) / a.stride - 1) * a.n + idx[w]) * 12;
#pragma unroll
                for (int j = 0; j < 12; ++j) dst[j] = x12[j];
            }
        }
    }
#pragma unroll
    for (int w = 0; w < NWIN; ++w) {
        if (!ok[w]) continue;
        if (a.x9T) {
#pragma unroll
            for (int j = 0; j < 9; ++j) a.x9T[idx[w] * 9 + j] = z[w][j];
        }
        if (a.lag_out) {
#pragma unroll
            for (int j = 0; j < 12; ++j) a.lag_out[idx[w] * 12 + j] = Z[w][j];
        }
    }
}

struct PincSeArgs {
    PincParams p;
    ThrMap m;
    const double* X;     // [rows][12]
    const double* U;     // [rows][8]
    double* partial;     // [gridDim.x][BROV_MAX_H]
    long long rows, nwin;
    int nH;
    int H[BROV_MAX_H];
    int carry_steps;     // 0 = every window starts from zero lag
    long long win0, row0;
    const double* carry_lag0;  // [8][3] lag state of the thruster-map object before window 0, or nullptr (zeros)
};

__global__ void __launch_bounds__(PB2, 1) pinc_se_kernel(const __grid_constant__ PincSeArgs a) {
    extern __shared__ __align__(16) float smf[];
    float* sw = smf;
    float* hbase = smf + NW + threadIdx.x;
    __shared__ double red[PB2 / 32][BROV_MAX_H];
    load_weights_n(a.p.w, sw, PB2);
    __syncthreads();
    double se[BROV_MAX_H];
#pragma unroll
    for (int h = 0; h < BROV_MAX_H; ++h) se[h] = 0.0;
    const int hmax = a.H[a.nH - 1];
    long long kr[NWIN];
    int nst[NWIN];
    float z[NWIN][NIN], xn[NWIN][9];
    double Z[NWIN][12];
#pragma unroll
    for (int w = 0; w < NWIN; ++w) {
        const long long gk = (long long)blockIdx.x * (NWIN * PB2) + w * PB2 + threadIdx.x;
        const bool live = gk < a.nwin;
        const long long k = live ? gk : a.nwin - 1;     // dead slots shadow the last window and score nothing
        kr[w] = k + (a.win0 - a.row0);
        const long long room = a.rows - 1 - kr[w];
        nst[w] = live ? (int)(room < hmax ? (room < 0 ? 0 : room) : hmax) : 0;
        double x12[12];
#pragma unroll
        for (int j = 0; j < 12; ++j) x12[j] = __ldg(a.X + kr[w] * 12 + j);
        x12_to_9(x12, z[w]);
        z[w][13] = (float)a.m.dt;
#pragma unroll
        for (int j = 0; j < 12; ++j) Z[w][j] = 0.0;
        if (a.carry_steps > 0 && live) {
            // the reference's single thruster-map object has seen windows 0..k-1, H steps each: replay the tail of that
            // history that is distinguishable from zero in floating point
            const long long H0 = a.H[0];
            const long long total = (a.win0 + k) * H0;
            const long long m = total < a.carry_steps ? total : a.carry_steps;
            if (a.carry_lag0 && total <= a.carry_steps) {  // the object's initial lag state is still visible
                double l24[24];
#pragma unroll
                for (int j = 0; j < 24; ++j) l24[j] = __ldg(a.carry_lag0 + j);
                project4(a.m, l24, Z[w]);
            }
            for (long long s = total - m; s < total; ++s) {
                const long long ww = s / H0;
                const long long row = ww + (s - ww * H0) - a.row0;
                double u8[8], u4[4];
#pragma unroll
                for (int j = 0; j < 8; ++j) u8[j] = __ldg(a.U + row * 8 + j);
                thruster_map4(a.m, u8, Z[w], u4);
            }
        }
    }
    const int nmax = nst[0] > nst[1] ? nst[0] : nst[1];
    for (int j = 0; j < nmax; ++j) {
#pragma unroll
        for (int w = 0; w < NWIN; ++w) {
            // a window that has run out of rows keeps stepping on its last valid input row; its results are ignored
            const long long row = kr[w] + (j < nst[w] ? j : (nst[w] > 0 ? nst[w] - 1 : 0));
            double u8[8], u4[4];
#pragma unroll
            for (int q = 0; q < 8; ++q) u8[q] = __ldg(a.U + row * 8 + q);
            thruster_map4(a.m, u8, Z[w], u4);
#pragma unroll
            for (int q = 0; q < 4; ++q) z[w][9 + q] = (float)u4[q];
        }
        pinc_forward<NWIN, PB2>(sw, a.p.beta, hbase, z, xn);
#pragma unroll
        for (int w = 0; w < NWIN; ++w) {
#pragma unroll
            for (int q = 0; q < 9; ++q) z[w][q] = xn[w][q];
#pragma unroll
            for (int h = 0; h < BROV_MAX_H; ++h) {
                if (h < a.nH && j + 1 == a.H[h] && j < nst[w]) {
                    double x12[12];
                    x9_to_12(xn[w], x12);
                    const double* tgt = a.X + (kr[w] + j + 1) * 12;
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 12; ++q) {
                        const double e = x12[q] - __ldg(tgt + q);
                        s += e * e;
                    }
                    se[h] += s;
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int h = 0; h < BROV_MAX_H; ++h) {
        double v = se[h];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][h] = v;
    }
    __syncthreads();
    if (threadIdx.x < BROV_MAX_H) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < PB2 / 32; ++w) v += red[w][threadIdx.x];
        a.partial[(long long)blockIdx.x * BROV_MAX_H + threadIdx.x] = v;
    }
}

__global__ void __launch_bounds__(256) pinc_finish_kernel(const double* __restrict__ partial, int nblocks,
                                                          double* __restrict__ out) {
    __shared__ double sh[256];
    for (int h = 0; h < BROV_MAX_H; ++h) {
        double v = 0.0;
        for (int b = threadIdx.x; b < nblocks; b += 256) v += partial[(long long)b * BROV_MAX_H + h];
        sh[threadIdx.x] = v;
        __syncthreads();
        for (int s = 128; s > 0; s >>= 1) {
            if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) out[h] = sh[0];
        __syncthreads();
    }
}

#include "brov_pinc_tc.cuh"

// PINcNet.forward on the tensor cores (brov_pinc_tc.cuh): persistent CTAs of three 128-row tiles, one thread per row
__global__ void __launch_bounds__(TC_THREADS, 1) pinc_forward_tc_kernel(const float* __restrict__ wtc,
                                                                         const __grid_constant__ TcAct act,
                                                                         const float* __restrict__ Zin,
                                                                         float* __restrict__ out, long long n) {
    extern __shared__ __align__(1024) float smf[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bars[TC_TILES];
    TcCtx c;
    tc_setup(c, smf, wtc, &tmem_slot, bars);
    const long long ntiles = (n + TC_M - 1) / TC_M;
    for (long long tile = (long long)blockIdx.x * TC_TILES + c.tile; tile < ntiles; tile += (long long)gridDim.x * TC_TILES) {
        const long long gi = tile * TC_M + c.row;
        const long long i = gi < n ? gi : n - 1;
        float z[NIN], xn[9];
#pragma unroll
        for (int j = 0; j < NIN; ++j) z[j] = Zin[i * NIN + j];
        tc_put_inputs(c, z, z[13], z + 9);
        pinc_net_tc(c, act);
        tc_layer_wait(c);
        tc_residual(c, z, xn);
        if (gi < n) {
#pragma unroll
            for (int j = 0; j < 9; ++j) out[i * 9 + j] = xn[j];
        }
    }
    tc_teardown(&tmem_slot);
}

// multistep_rmse_endpoint_pinc on the tensor cores: a tile = 128 consecutive windows stepped together through the
// longest horizon; a thread carries its window's network state and projected lag states, scores the endpoints, and
// evaluates the fp64 thruster map of the NEXT step while the output layer's MMAs of the current one run.
__global__ void __launch_bounds__(TC_THREADS, 1) pinc_se_tc_kernel(const __grid_constant__ PincSeArgs a,
                                                                    const float* __restrict__ wtc,
                                                                    const __grid_constant__ TcAct act) {
    extern __shared__ __align__(1024) float smf[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bars[TC_TILES];
    __shared__ double red[TC_THREADS / 32][BROV_MAX_H];
    TcCtx c;
    tc_setup(c, smf, wtc, &tmem_slot, bars);
    double se[BROV_MAX_H];
#pragma unroll
    for (int h = 0; h < BROV_MAX_H; ++h) se[h] = 0.0;
    const int hmax = a.H[a.nH - 1];
    const float dtf = (float)a.m.dt;
    const long long ntiles = (a.nwin + TC_M - 1) / TC_M;
    // the CTA's three tiles walk the tile list independently (tile-wide barriers only inside the loop)
    for (long long tile = (long long)blockIdx.x * TC_TILES + c.tile; tile < ntiles; tile += (long long)gridDim.x * TC_TILES) {
        const long long gk = tile * TC_M + c.row;
        const bool live = gk < a.nwin;
        const long long k = live ? gk : a.nwin - 1;     // dead slots shadow the last window and score nothing
        const long long kr = k + (a.win0 - a.row0);
        const long long room = a.rows - 1 - kr;
        const int nst = live ? (int)(room < hmax ? (room < 0 ? 0 : room) : hmax) : 0;
        float z[9], u4f[4];
        double Z[12];
        // a window that has run out of rows keeps stepping on its last valid input row; its results are ignored
        auto thrust = [&](int j) {
            const long long row = kr + (j < nst ? j : (nst > 0 ? nst - 1 : 0));
            double u8[8], u4[4];
#pragma unroll
            for (int q = 0; q < 8; ++q) u8[q] = __ldg(a.U + row * 8 + q);
            thruster_map4(a.m, u8, Z, u4);
#pragma unroll
            for (int q = 0; q < 4; ++q) u4f[q] = (float)u4[q];
        };
        {
            double x12[12];
#pragma unroll
            for (int j = 0; j < 12; ++j) x12[j] = __ldg(a.X + kr * 12 + j);
            x12_to_9(x12, z);
        }
#pragma unroll
        for (int j = 0; j < 12; ++j) Z[j] = 0.0;
        if (a.carry_steps > 0 && live) {
            const long long H0 = a.H[0];
            const long long total = (a.win0 + k) * H0;
            const long long m = total < a.carry_steps ? total : a.carry_steps;
            if (a.carry_lag0 && total <= a.carry_steps) {
                double l24[24];
#pragma unroll
                for (int j = 0; j < 24; ++j) l24[j] = __ldg(a.carry_lag0 + j);
                project4(a.m, l24, Z);
            }
            for (long long s = total - m; s < total; ++s) {
                const long long ww = s / H0;
                const long long row = ww + (s - ww * H0) - a.row0;
                double u8[8], u4[4];
#pragma unroll
                for (int j = 0; j < 8; ++j) u8[j] = __ldg(a.U + row * 8 + j);
                thruster_map4(a.m, u8, Z, u4);
            }
        }
        thrust(0);
        for (int j = 0; j < hmax; ++j) {     // uniform trip count: the network holds tile-wide barriers
            tc_put_inputs(c, z, dtf, u4f);
            pinc_net_tc(c, act);
            if (j + 1 < hmax) thrust(j + 1);       // the fp64 thruster map of the next step under the output layer's MMAs
            tc_layer_wait(c);
            float xn[9];
            tc_residual(c, z, xn);
#pragma unroll
            for (int q = 0; q < 9; ++q) z[q] = xn[q];
#pragma unroll
            for (int h = 0; h < BROV_MAX_H; ++h) {
                if (h < a.nH && j + 1 == a.H[h] && j < nst) {
                    double x12[12];
                    x9_to_12(xn, x12);
                    const double* tgt = a.X + (kr + j + 1) * 12;
                    double s = 0.0;
#pragma unroll
                    for (int q = 0; q < 12; ++q) {
                        const double e = x12[q] - __ldg(tgt + q);
                        s += e * e;
                    }
                    se[h] += s;
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int h = 0; h < BROV_MAX_H; ++h) {
        double v = se[h];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_down_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][h] = v;
    }
    __syncthreads();
    if (threadIdx.x < BROV_MAX_H) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < TC_THREADS / 32; ++w) v += red[w][threadIdx.x];
        a.partial[(long long)blockIdx.x * BROV_MAX_H + threadIdx.x] = v;
    }
    tc_teardown(&tmem_slot);
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------------
struct brov_pinc {
    int device;
    int num_sms;
    bool use_tc;     // dense layers on the tensor cores (default; BROV_PINC_TC=0 selects the CUDA-core kernels)
    TcAct act;       // folded activation / LayerNorm scalars of the tensor-core path
    float* wtc;      // tensor-core blob (TC_NW floats): hi / lo TF32 parts of every layer in UMMA core-matrix layout
    float* w;        // packed blob on the device
    float beta[4];
    ThrMap m;
    bool have_map;
    double* partial;
    size_t cap_partial;
    bool attr_set;
};

static int pinc_attrs(brov_pinc* h) {
    if (h->attr_set) return BROV_OK;
    BROV_CUDA_TRY(cudaFuncSetAttribute(pinc_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    BROV_CUDA_TRY(cudaFuncSetAttribute(pinc_rollout_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES2));
    BROV_CUDA_TRY(cudaFuncSetAttribute(pinc_se_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES2));
    BROV_CUDA_TRY(cudaFuncSetAttribute(pinc_forward_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    BROV_CUDA_TRY(cudaFuncSetAttribute(pinc_se_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TC_SMEM_BYTES));
    h->attr_set = true;
    return BROV_OK;
}

extern "C" int brov_pinc_create(int device, const brov_pinc_weights* wts, brov_pinc_t** out) {
    if (!out) return brov::fail_msg(BROV_EINVAL, "out is NULL");
    *out = nullptr;
    if (!wts || wts->struct_size != sizeof(brov_pinc_weights)) return brov::fail_msg(BROV_EINVAL, "brov_pinc_weights size mismatch (ABI %d)", BROV_ABI_VERSION);
    if (wts->n_hidden_layers != 4 || wts->hidden != HID) return brov::fail_msg(BROV_EUNSUPPORTED, "compiled for the reference's 4 x 64 network (PINc_HIDDEN), got %d x %d", wts->n_hidden_layers, wts->hidden);
    for (int l = 0; l < 5; ++l)
        if (!wts->W[l] || !wts->b[l] || (l < 4 && (!wts->ln_w[l] || !wts->ln_b[l]))) return brov::fail_msg(BROV_EINVAL, "weights of layer %d are NULL", l);
    int ndev = 0;
    BROV_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return brov::fail_msg(BROV_EINVAL, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    BROV_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return brov::fail_msg(BROV_EUNSUPPORTED, "libbrov is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    BROV_CUDA_TRY(cudaSetDevice(device));
    // pack: linear layers transposed to input-major so that one input's fan-out is a contiguous row
    float* blob = new (std::nothrow) float[NW];
    brov_pinc* h = new (std::nothrow) brov_pinc();
    if (!blob || !h) { delete[] blob; delete h; return brov::fail_msg(BROV_ENOMEM, "out of host memory"); }
    memset(blob, 0, NW * sizeof(float));
    for (int i = 0; i < NIN; ++i) for (int j = 0; j < HID; ++j) blob[OFF_W0 + i * HID + j] = wts->W[0][j * NIN + i];
    for (int j = 0; j < HID; ++j) {
        blob[OFF_B0 + j] = wts->b[0][j];
        blob[OFF_B0 + HID + j] = wts->ln_w[0][j];
        blob[OFF_B0 + 2 * HID + j] = wts->ln_b[0][j];
    }
    for (int l = 0; l < 3; ++l) {
        float* L = blob + OFF_L1 + l * LSTRIDE;
        for (int i = 0; i < HID; ++i) for (int j = 0; j < HID; ++j) L[i * HID + j] = wts->W[l + 1][j * HID + i];
        for (int j = 0; j < HID; ++j) {
            L[HID * HID + j] = wts->b[l + 1][j];
            L[HID * HID + HID + j] = wts->ln_w[l + 1][j];
            L[HID * HID + 2 * HID + j] = wts->ln_b[l + 1][j];
        }
    }
    for (int i = 0; i < HID; ++i) for (int j = 0; j < 9; ++j) blob[OFF_W4 + i * 12 + j] = wts->W[4][j * HID + i];
    for (int j = 0; j < 9; ++j) blob[OFF_B4 + j] = wts->b[4][j];
    memset(h, 0, sizeof(*h));
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    {
        const char* env = getenv("BROV_PINC_TC");
        h->use_tc = !(env && env[0] == '0');
    }
    for (int l = 0; l < 4; ++l) h->beta[l] = wts->beta[l];
    // tensor-core blob: every weight split into two TF32 numbers (round to nearest), W[n][k] at tc_off(rows, n, k)
    float* tcb = new (std::nothrow) float[TC_NW];    // TF32 parts, biases, then FP16 copies of the high parts
    if (!tcb) { delete[] blob; delete h; return brov::fail_msg(BROV_ENOMEM, "out of host memory"); }
    memset(tcb, 0, TC_NW * sizeof(float));
    unsigned short* tcb16 = reinterpret_cast<unsigned short*>(tcb);
    auto rna = [](float x) {
        uint32_t b;
        memcpy(&b, &x, 4);
        if ((b & 0x7f800000u) != 0x7f800000u) b = (b + 0x1000u) & 0xffffe000u;
        float r;
        memcpy(&r, &b, 4);
        return r;
    };
    // x (double) -> two TF32 numbers, x = hi + lo to 2^-22 relative
    auto put = [&](int hi_off, int lo_off, int h16_off, int rows, int n, int k, double w) {
        const float hi = rna((float)w);
        tcb[hi_off + tc_off(rows, n, k)] = hi;
        tcb[lo_off + tc_off(rows, n, k)] = rna((float)(w - (double)hi));
        const __half w16 = __float2half_rn((float)w);      // multiplies the activations' FP16 low parts
        memcpy(&tcb16[2 * h16_off + tc_off16(rows, n, k)], &w16, 2);
    };
    // Folds (double precision): LayerNorm l's affine into layer l + 1 (W' = W diag(ln_w), b' = b + W ln_b); the
    // activation's input scale beta log2(e) into the hidden layers' biases; its output scale ln 2 / (beta + 1e-12) into
    // LayerNorm's epsilon (see brov_pinc_tc.cuh).
    const double LOG2E = 1.4426950408889634, LN2 = 0.6931471805599453;
    for (int l = 0; l < 4; ++l) {
        const double kappa = LN2 / ((double)wts->beta[l] + 1e-12);
        h->act.sc[l] = (float)((double)wts->beta[l] * LOG2E);
        h->act.eps[l] = (float)(1e-5 / (kappa * kappa));
        h->act.sg[l] = kappa < 0.0 ? -1.0f : 1.0f;
    }
    for (int l = 0; l <= 4; ++l) {
        const int K = l == 0 ? NIN : HID, N = l == 4 ? 9 : HID;
        const int hi_off = l == 0 ? TC_L0_HI : (l == 4 ? TC_L4_HI : TC_L1 + (l - 1) * 8192);
        const int lo_off = l == 0 ? TC_L0_LO : (l == 4 ? TC_L4_LO : TC_L1 + (l - 1) * 8192 + 4096);
        const int rows = l == 4 ? 16 : HID;
        const int h16_off = l == 0 ? TC_H16_L0 : (l == 4 ? TC_H16_L4 : TC_H16_L1 + (l - 1) * 2048);
        for (int n = 0; n < N; ++n) {
            double bias = wts->b[l][n];
            if (l == 0) {
                for (int k = 0; k < 16; ++k)       // layer 0's K axis in the order the kernels store it (tc_kmap0)
                    if (tc_kmap0(k) >= 0) put(hi_off, lo_off, h16_off, rows, n, k, wts->W[0][n * NIN + tc_kmap0(k)]);
            } else {
                for (int k = 0; k < K; ++k) {
                    const double w = wts->W[l][n * K + k];
                    put(hi_off, lo_off, h16_off, rows, n, k, w * (double)wts->ln_w[l - 1][k]);
                    bias += w * (double)wts->ln_b[l - 1][k];
                }
            }
            tcb[TC_PAR + l * 64 + n] = (float)(l < 4 ? bias * (double)wts->beta[l] * LOG2E : bias);
        }
    }
    cudaError_t e = cudaMalloc(&h->w, NW * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->w, blob, NW * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMalloc(&h->wtc, TC_NW * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->wtc, tcb, TC_NW * sizeof(float), cudaMemcpyHostToDevice);
    delete[] blob;
    delete[] tcb;
    if (e != cudaSuccess) {
        cudaFree(h->w);
        cudaFree(h->wtc);
        delete h;
        return brov::fail_msg(BROV_ECUDA, "brov_pinc_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return BROV_OK;
}

extern "C" void brov_pinc_destroy(brov_pinc_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->w);
    cudaFree(h->wtc);
    cudaFree(h->partial);
    delete h;
}

extern "C" int brov_pinc_set_thruster_map(brov_pinc_t* h, double dt, const double* Ad, const double* Bd,
                                          const double* alloc) {
    if (!h || !Ad || !Bd || !alloc) return brov::fail_msg(BROV_EINVAL, "NULL argument");
    if (!(dt > 0.0)) return brov::fail_msg(BROV_EINVAL, "dt must be > 0");
    memcpy(h->m.Ad, Ad, sizeof(h->m.Ad));
    memcpy(h->m.Bd, Bd, sizeof(h->m.Bd));
    h->m.Cc[0] = 0.0; h->m.Cc[1] = 5.992; h->m.Cc[2] = 3.317;   // fossen/BlueROV2.py:480
    const int rows[4] = {0, 1, 2, 5};
    for (int r = 0; r < 4; ++r)
        for (int i = 0; i < 8; ++i) {
            const bool nz = (r < 2 || r == 3) ? (i < 4) : (i >= 4);
            const double v = alloc[rows[r] * 8 + i];
            if (!nz && v != 0.0) return brov::fail_msg(BROV_EUNSUPPORTED, "allocation[%d][%d] must be zero (BlueROV2 heavy layout)", rows[r], i);
            h->m.a4[r][i] = v;
        }
    h->m.dt = dt;
    h->have_map = true;
    return BROV_OK;
}

extern "C" int brov_pinc_forward(brov_pinc_t* h, const float* z_dev, float* out_dev, long long n, void* stream) {
    if (!h || (n > 0 && (!z_dev || !out_dev))) return brov::fail_msg(BROV_EINVAL, "NULL argument");
    if (n <= 0) return BROV_OK;
    BROV_CUDA_TRY(cudaSetDevice(h->device));
    int rc = pinc_attrs(h);
    if (rc) return rc;
    PincParams p;
    p.w = h->w;
    memcpy(p.beta, h->beta, sizeof(p.beta));
    if (h->use_tc) {
        const long long ctas = ((n + TC_M - 1) / TC_M + TC_TILES - 1) / TC_TILES;
        const unsigned grid = (unsigned)(ctas < h->num_sms ? ctas : h->num_sms);
        pinc_forward_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(
            h->wtc, h->act, z_dev, out_dev, n);
    } else {
        pinc_forward_kernel<<<(unsigned)((n + PB - 1) / PB), PB, SMEM_BYTES, (cudaStream_t)stream>>>(p, z_dev, out_dev, n);
    }
    BROV_CUDA_TRY(cudaGetLastError());
    return BROV_OK;
}

extern "C" int brov_pinc_rollout(brov_pinc_t* h, const brov_pinc_rollout_desc* d, void* stream) {
    if (!h) return brov::fail_msg(BROV_EINVAL, "NULL handle");
    if (!d || d->struct_size != sizeof(brov_pinc_rollout_desc)) return brov::fail_msg(BROV_EINVAL, "brov_pinc_rollout_desc size mismatch (ABI %d)", BROV_ABI_VERSION);
    if (!h->have_map) return brov::fail_msg(BROV_EINVAL, "call brov_pinc_set_thruster_map first");
    if (d->n < 0 || d->steps < 0 || d->steps > 0x7fffffffLL) return brov::fail_msg(BROV_EINVAL, "n / steps out of range");
    if (d->n == 0) return BROV_OK;
    if (!d->x0_dev || (d->steps > 0 && !d->u_dev)) return brov::fail_msg(BROV_EINVAL, "NULL array");
    if (d->traj_dev && d->stride < 1) return brov::fail_msg(BROV_EINVAL, "stride must be >= 1 with a trajectory buffer");
    BROV_CUDA_TRY(cudaSetDevice(h->device));
    int rc = pinc_attrs(h);
    if (rc) return rc;
    PincRollArgs a;
    a.p.w = h->w;
    memcpy(a.p.beta, h->beta, sizeof(a.p.beta));
    a.m = h->m;
    a.x0 = (const double*)d->x0_dev; a.U = (const double*)d->u_dev;
    a.u_stride_t = d->u_stride_t; a.u_stride_n = d->u_stride_n;
    a.lag_in = (const double*)d->lag_in_dev; a.lag_out = (double*)d->lag_out_dev;
    a.traj = (double*)d->traj_dev; a.x9T = (float*)d->x9T_dev;
    a.n = d->n; a.steps = (int)d->steps; a.stride = d->traj_dev ? (int)d->stride : 1;
    pinc_rollout_kernel<<<(unsigned)((d->n + NWIN * PB2 - 1) / (NWIN * PB2)), PB2, SMEM_BYTES2, (cudaStream_t)stream>>>(a);
    BROV_CUDA_TRY(cudaGetLastError());
    return BROV_OK;
}

extern "C" int brov_pinc_multistep_se(brov_pinc_t* h, const brov_pinc_se_desc* d, void* stream) {
    if (!h) return brov::fail_msg(BROV_EINVAL, "NULL handle");
    if (!d || d->struct_size != sizeof(brov_pinc_se_desc)) return brov::fail_msg(BROV_EINVAL, "brov_pinc_se_desc size mismatch (ABI %d)", BROV_ABI_VERSION);
    if (!h->have_map) return brov::fail_msg(BROV_EINVAL, "call brov_pinc_set_thruster_map first");
    if (d->n_horizons < 1 || d->n_horizons > BROV_MAX_H) return brov::fail_msg(BROV_EINVAL, "n_horizons must be 1..%d", BROV_MAX_H);
    for (int q = 0; q < d->n_horizons; ++q)
        if (d->horizons[q] < 1 || (q && d->horizons[q] <= d->horizons[q - 1])) return brov::fail_msg(BROV_EINVAL, "horizons must be >= 1 and strictly ascending");
    if (d->carry_steps < 0 || (d->carry_steps > 0 && d->n_horizons != 1)) return brov::fail_msg(BROV_EINVAL, "a carried lag scores one horizon per call");
    if (d->window0 < 0 || d->row0 < 0 || d->row0 > d->window0) return brov::fail_msg(BROV_EINVAL, "window0 / row0 out of range");
    if (d->rows < 0 || d->n_windows < 0 || d->n_windows > d->rows) return brov::fail_msg(BROV_EINVAL, "rows / n_windows out of range");
    // every window's first row must be local: window k starts at local row k + (window0 - row0)
    if (d->n_windows + (d->window0 - d->row0) > d->rows)
        return brov::fail_msg(BROV_EINVAL, "n_windows = %lld starting at local row %lld exceed the %lld rows of the series", d->n_windows, d->window0 - d->row0, d->rows);
    if (d->carry_steps == 0 && d->window0 != d->row0) return brov::fail_msg(BROV_EINVAL, "rows before the first window are only meaningful with a carried lag");
    if (d->carry_steps > 0) {
        // the replay of window0's history reads input rows back to window (window0 * H - carry_steps) / H
        const long long H = d->horizons[0];
        const long long first_hist = d->window0 * H - d->carry_steps;
        const long long need_row = first_hist > 0 ? first_hist / H : 0;
        if (d->row0 > need_row)
            return brov::fail_msg(BROV_EINVAL, "carried lag: the shard must start at row %lld or earlier (row0 = %lld)", need_row, d->row0);
    }
    if (!d->se_out_dev || (d->n_windows > 0 && (!d->X_dev || !d->U_dev))) return brov::fail_msg(BROV_EINVAL, "NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    BROV_CUDA_TRY(cudaSetDevice(h->device));
    if (d->n_windows == 0) {
        BROV_CUDA_TRY(cudaMemsetAsync(d->se_out_dev, 0, BROV_MAX_H * sizeof(double), st));
        return BROV_OK;
    }
    int rc = pinc_attrs(h);
    if (rc) return rc;
    const size_t per_cta = TC_TILES;
    const size_t tiles = ((size_t)((d->n_windows + TC_M - 1) / TC_M) + per_cta - 1) / per_cta;   // CTAs of per_cta tiles
    const size_t nblocks = h->use_tc ? (tiles < (size_t)h->num_sms ? tiles : (size_t)h->num_sms)
                                     : (size_t)((d->n_windows + NWIN * PB2 - 1) / (NWIN * PB2));
    if (nblocks * BROV_MAX_H > h->cap_partial) {
        cudaFree(h->partial);
        h->partial = nullptr; h->cap_partial = 0;
        BROV_CUDA_TRY(cudaMalloc(&h->partial, nblocks * BROV_MAX_H * sizeof(double)));
        h->cap_partial = nblocks * BROV_MAX_H;
    }
    PincSeArgs a;
    a.p.w = h->w;
    memcpy(a.p.beta, h->beta, sizeof(a.p.beta));
    a.m = h->m;
    a.X = (const double*)d->X_dev; a.U = (const double*)d->U_dev; a.partial = h->partial;
    a.rows = d->rows; a.nwin = d->n_windows; a.nH = d->n_horizons;
    for (int q = 0; q < BROV_MAX_H; ++q) a.H[q] = q < d->n_horizons ? d->horizons[q] : 0x7fffffff;
    a.carry_steps = d->carry_steps; a.win0 = d->window0; a.row0 = d->row0;
    a.carry_lag0 = (const double*)d->carry_lag0_dev;
    if (h->use_tc) pinc_se_tc_kernel<<<(unsigned)nblocks, TC_THREADS, TC_SMEM_BYTES, st>>>(a, h->wtc, h->act);
    else pinc_se_kernel<<<(unsigned)nblocks, PB2, SMEM_BYTES2, st>>>(a);
    pinc_finish_kernel<<<1, 256, 0, st>>>(h->partial, (int)nblocks, d->se_out_dev);
    BROV_CUDA_TRY(cudaGetLastError());
    return BROV_OK;
}
