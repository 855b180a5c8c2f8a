// brov_koopman.cu — scoring and simulation of a fitted Koopman EDMDc model on sm_100a (float64, as the reference).
//
// Reference (ViktorNfa/bluerov2_dynamics, Koopman/koopmanEDMDc.py):
//   _rbf_mat / _lift    :41-49, 221-238   phi(x) = [x, exp(-gamma (|x|^2 + |c_j|^2 - 2 x.c_j))], d = n + k
//   evaluate            :157-170          one-step RMSE  (= multistep_rmse with H = 1)
//   multistep_rmse      :172-200          Z <- Z A^T + U_t B^T for t < H on all windows at once, decode = first n
//   simulate            :202-216          z <- A z + B u, record the first n coordinates
//
// B200-first formulation.  The reference pushes the full lifted state (d = 512) of every window through H dense
// d x d products: 2 d^2 H flop per window.  Only the first n = 12 coordinates of the result are ever looked at, so
// the engine propagates the n ROWS of the decoder through the model instead, once per model:
//       W_t = E A^t  (n x d),   G_j = W_j B  (n x r),   E = [I_n 0]
//       x_hat_k = W_H phi(x_k) + sum_{t<H} G_{H-1-t} u_{k+t}
// which is the same linear map evaluated in a different association order (agreement with the sequential form:
// 1e-15 relative on the RMSE, tests/test_gpu_compare.py) at 2 n (d + r H) flop per window — a 1600-fold reduction
// for d = 512, H = 100.  What remains per window is the RBF lift (k exponentials) with an n x d mat-vec per horizon
// (koop_liftw_kernel: one thread per window, centres and decoder rows broadcast from shared memory, evaluated ONCE for
// all requested horizons) and an FIR filter over the inputs per horizon (koop_fir_se_kernel: four windows per thread,
// taps and a transposed input tile in shared memory).  FP64-pipe bound.
#include <cmath>
#include <cstring>
#include <new>
#include <vector>

#include "brov_internal.cuh"

namespace {

constexpr int KB = 128;        // threads per block of the lift kernel
constexpr int NMAX = 16;       // state dimension bound (registers)
constexpr int RMAX = 8;        // input dimension bound

// ---------------------------------------------------------------------------------------------------------------
// W_{t+1} = W_t A  (n x d times d x d), one launch per power: a sequential chain, so each launch has to be short.
// Block: 16 output columns x 16 slices of the inner dimension; W_t is staged in shared memory, the A loads of a
// thread are independent and unrolled so that several are in flight; fixed-order reduction over the slices.
// ---------------------------------------------------------------------------------------------------------------
constexpr int PW_COLS = 16, PW_SLICES = 16;
__global__ void __launch_bounds__(PW_COLS * PW_SLICES) koop_power_kernel(const double* __restrict__ W,
                                                                          const double* __restrict__ A,
                                                                          double* __restrict__ Wn, int n, int d) {
    extern __shared__ double pw_sm[];
    double* sWt = pw_sm;                                   // [d][NMAX]: W_t transposed, zero-padded rows n..NMAX
    double* part = pw_sm + (size_t)d * NMAX;               // [PW_SLICES][NMAX][PW_COLS + 1]
    const int cx = threadIdx.x % PW_COLS, sl = threadIdx.x / PW_COLS;
    const int c = blockIdx.x * PW_COLS + cx;
    for (int e = threadIdx.x; e < d * NMAX; e += PW_COLS * PW_SLICES) {
        const int q = e / NMAX, i = e - q * NMAX;
        sWt[e] = i < n ? W[(size_t)i * d + q] : 0.0;
    }
    __syncthreads();
    double acc[NMAX];
#pragma unroll
    for (int i = 0; i < NMAX; ++i) acc[i] = 0.0;
    if (c < d) {
#pragma unroll 4
        for (int q = sl; q < d; q += PW_SLICES) {
            const double a = __ldg(A + (size_t)q * d + c);
            const double* w = sWt + (size_t)q * NMAX;
#pragma unroll
            for (int i = 0; i < NMAX; ++i) acc[i] = fma(w[i], a, acc[i]);
        }
    }
#pragma unroll
    for (int i = 0; i < NMAX; ++i) part[(sl * NMAX + i) * (PW_COLS + 1) + cx] = acc[i];
    __syncthreads();
    for (int e = threadIdx.x; e < n * PW_COLS; e += PW_COLS * PW_SLICES) {
        const int i = e / PW_COLS, x = e - i * PW_COLS;
        const int cc = blockIdx.x * PW_COLS + x;
        if (cc >= d) continue;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < PW_SLICES; ++k) s += part[(k * NMAX + i) * (PW_COLS + 1) + x];
        Wn[(size_t)i * d + cc] = s;
    }
}

// G_j = W_j B for j in [j0, j1): block per j, thread per (i, c)
__global__ void koop_gain_kernel(const double* __restrict__ W, const double* __restrict__ B, double* __restrict__ G,
                                 int n, int r, int d, int j0) {
    const int j = j0 + blockIdx.x;
    const double* Wj = W + (size_t)j * n * d;
    for (int e = threadIdx.x; e < n * r; e += blockDim.x) {
        const int i = e / r, c = e - i * r;
        double s = 0.0;
        for (int q = 0; q < d; ++q) s = fma(Wj[(size_t)i * d + q], B[(size_t)q * r + c], s);
        G[(size_t)j * n * r + e] = s;
    }
}

__global__ void koop_c2_kernel(const double* __restrict__ C, double* __restrict__ c2, int n, int k) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += C[(size_t)j * n + i] * C[(size_t)j * n + i];
    c2[j] = s;
}

// ---------------------------------------------------------------------------------------------------------------
// lift: Z[row] = [x, rbf_1(x) .. rbf_k(x)]
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(KB) koop_lift_kernel(const double* __restrict__ X, const double* __restrict__ C,
                                                       const double* __restrict__ c2, double gamma, long long rows,
                                                       int n, int k, double* __restrict__ Z) {
    extern __shared__ double sm[];
    double* sC = sm;            // [k][n]
    double* sc2 = sm + (size_t)k * n;
    for (int e = threadIdx.x; e < k * n; e += KB) sC[e] = C[e];
    for (int e = threadIdx.x; e < k; e += KB) sc2[e] = c2[e];
    __syncthreads();
    const long long row = (long long)blockIdx.x * KB + threadIdx.x;
    if (row >= rows) return;
    double x[NMAX], x2 = 0.0;
#pragma unroll
    for (int i = 0; i < NMAX; ++i) {
        x[i] = i < n ? X[row * n + i] : 0.0;
        x2 = fma(x[i], x[i], x2);
    }
    const int d = n + k;
    double* z = Z + row * d;
    for (int i = 0; i < n; ++i) z[i] = x[i];
    for (int j = 0; j < k; ++j) {
        double dot = 0.0;
#pragma unroll
        for (int i = 0; i < NMAX; ++i)
            if (i < n) dot = fma(x[i], sC[j * n + i], dot);
        z[n + j] = exp(-gamma * (x2 + sc2[j] - 2.0 * dot));
    }
}

// ---------------------------------------------------------------------------------------------------------------
// Several horizons in one pass.  The lift phi(x_k) (500 exponentials) does not depend on the horizon, so it is evaluated
// ONCE per window and pushed through the decoder rows of every requested horizon:
//   koop_liftw_kernel   P[h][k][:] = W_{H_h} phi(x_k)                     one thread per window, persistent blocks
//   koop_fir_se_kernel  x_hat = P[h][k] + sum_t G_{H-1-t} u_{k+t}, squared error vs X[k+H]   four windows per thread
// ---------------------------------------------------------------------------------------------------------------
constexpr int LWT = 256;       // threads per block of the lift kernel
constexpr int LWW = 2;         // windows per thread: every centre / decoder column read from shared memory feeds two windows

template <int N, int NH>
__global__ void __launch_bounds__(LWT, 1) koop_liftw_kernel(const double* __restrict__ X, const double* __restrict__ C,
                                                           const double* __restrict__ c2, const double* __restrict__ W,
                                                           const int* __restrict__ Hs, double gamma, long long nwin,
                                                           int k, double* __restrict__ P) {
    extern __shared__ double sm[];
    const int d = N + k;
    double* sC = sm;                          // [k][N]
    double* sc2 = sC + (size_t)k * N;         // [k]
    double* sW = sc2 + k;                     // [NH][d][N]  (transposed decoder rows of each horizon)
    for (int e = threadIdx.x; e < k * N; e += LWT) sC[e] = C[e];
    for (int e = threadIdx.x; e < k; e += LWT) sc2[e] = c2[e];
    for (int h = 0; h < NH; ++h) {
        const double* WH = W + (size_t)Hs[h] * N * d;
        for (int e = threadIdx.x; e < d * N; e += LWT) {
            const int q = e / N, i = e - q * N;
            sW[(size_t)h * d * N + e] = WH[(size_t)i * d + q];
        }
    }
    __syncthreads();
    const long long ntiles = (nwin + LWT * LWW - 1) / (LWT * LWW);
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        long long v[LWW];
        bool ok[LWW];
        double x[LWW][N], acc[LWW][NH][N], x2[LWW];
#pragma unroll
        for (int j = 0; j < LWW; ++j) {
            const long long w = tile * (LWT * LWW) + (long long)j * LWT + threadIdx.x;
            ok[j] = w < nwin;
            v[j] = ok[j] ? w : nwin - 1;
            x2[j] = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                x[j][i] = __ldg(X + v[j] * N + i);
                x2[j] = fma(x[j][i], x[j][i], x2[j]);
            }
#pragma unroll
            for (int h = 0; h < NH; ++h)
#pragma unroll
                for (int i = 0; i < N; ++i) acc[j][h][i] = 0.0;
        }
#pragma unroll
        for (int q = 0; q < N; ++q)
#pragma unroll
            for (int h = 0; h < NH; ++h)
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double wv = sW[((size_t)h * d + q) * N + i];
#pragma unroll
                    for (int j = 0; j < LWW; ++j) acc[j][h][i] = fma(wv, x[j][q], acc[j][h][i]);
                }
        const double mg = -gamma;
#pragma unroll 1
        for (int c = 0; c < k; ++c) {
            double dot[LWW];
#pragma unroll
            for (int j = 0; j < LWW; ++j) dot[j] = 0.0;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const double cv = sC[c * N + i];
#pragma unroll
                for (int j = 0; j < LWW; ++j) dot[j] = fma(x[j][i], cv, dot[j]);
            }
            const double cc = sc2[c];
            double e[LWW];
#pragma unroll
            for (int j = 0; j < LWW; ++j) e[j] = exp(mg * (x2[j] + cc - 2.0 * dot[j]));
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const double* wc = sW + ((size_t)h * d + N + c) * N;
#pragma unroll
                for (int i = 0; i < N; ++i) {
                    const double wv = wc[i];
#pragma unroll
                    for (int j = 0; j < LWW; ++j) acc[j][h][i] = fma(wv, e[j], acc[j][h][i]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < LWW; ++j) {
            if (!ok[j]) continue;
#pragma unroll
            for (int h = 0; h < NH; ++h)
#pragma unroll
                for (int i = 0; i < N; ++i) P[((size_t)h * nwin + v[j]) * N + i] = acc[j][h][i];
        }
    }
}

constexpr int FT = 256, FW = 4;   // threads per block and windows per thread of the FIR kernel (1024-window tiles)

// Window w at tap t reads input row w + t: a warp's 32 rows are 64 B apart, so a 128-bit global load touches 16 cache
// lines and the L1 data pipe saturates (93 % of peak at 27 % of the FP64 pipe in profiles/r01k_koopman_multi_raw.csv).
// The tile's rows [w0, w0 + FT FW + H) are therefore staged once into shared memory TRANSPOSED ([channel][row]), where
// the same access is a conflict-free 64-bit load of consecutive words.
template <int N, int R>
__global__ void __launch_bounds__(FT) koop_fir_se_kernel(const double* __restrict__ X, const double* __restrict__ U,
                                                         const double* __restrict__ P, const double* __restrict__ Gg,
                                                         int H, long long nwin, long long rows, int g_in_smem,
                                                         int u_in_smem, double* __restrict__ partial) {
    extern __shared__ double sm[];
    double* sG = sm;
    const int TR = FT * FW + H;                      // rows of the input tile
    double* sU = sm + (g_in_smem ? (size_t)H * N * R : 0);   // [R][TR]
    const long long w0 = (long long)blockIdx.x * (FT * FW);
    if (g_in_smem)
        for (int e = threadIdx.x; e < H * N * R; e += FT) sG[e] = Gg[e];
    if (u_in_smem) {
        for (int e = threadIdx.x; e < TR * R; e += FT) {
            const int row = e / R, c = e - row * R;
            const long long gr = w0 + row;
            sU[(size_t)c * TR + row] = gr < rows ? __ldg(U + gr * R + c) : 0.0;
        }
    }
    __syncthreads();
    const double* G = g_in_smem ? sG : Gg;
    __shared__ double red[FT / 32];
    long long v[FW];
    bool ok[FW];
    double acc[FW][N];
#pragma unroll
    for (int j = 0; j < FW; ++j) {
        const long long w = w0 + (long long)j * FT + threadIdx.x;
        ok[j] = w < nwin;
        v[j] = ok[j] ? w : nwin - 1;
#pragma unroll
        for (int i = 0; i < N; ++i) acc[j][i] = __ldg(P + v[j] * N + i);
    }
    for (int t = 0; t < H; ++t) {
        double u[FW][R];
        if (u_in_smem) {
#pragma unroll
            for (int j = 0; j < FW; ++j)
#pragma unroll
                for (int c = 0; c < R; ++c) u[j][c] = sU[(size_t)c * TR + (v[j] - w0) + t];
        } else {
#pragma unroll
            for (int j = 0; j < FW; ++j)
#pragma unroll
                for (int c = 0; c < R; ++c) u[j][c] = __ldg(U + (v[j] + t) * R + c);
        }
        const double* g = G + (size_t)(H - 1 - t) * N * R;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int c = 0; c < R; ++c) {
                const double gv = g[i * R + c];
#pragma unroll
                for (int j = 0; j < FW; ++j) acc[j][i] = fma(gv, u[j][c], acc[j][i]);
            }
    }
    double se = 0.0;
#pragma unroll
    for (int j = 0; j < FW; ++j) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double e = __ldg(X + (v[j] + H) * N + i) - acc[j][i];
            s = fma(e, e, s);
        }
        se += ok[j] ? s : 0.0;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) se += __shfl_down_sync(0xffffffffu, se, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = se;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
#pragma unroll
        for (int q = 0; q < FT / 32; ++q) s += red[q];
        partial[blockIdx.x] = s;
    }
}

__global__ void __launch_bounds__(256) koop_finish_kernel(const double* __restrict__ partial, int nblocks,
                                                          double* __restrict__ out) {
    __shared__ double sh[256];
    double v = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += 256) v += partial[b];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x == 0) out[0] = sh[0];
}

// ---------------------------------------------------------------------------------------------------------------
// simulate: X_pred[t][b] = W_t z0_b + sum_{s<t} G_{t-1-s} u_s   for t = 1..T; block per t, warp per output
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) koop_sim_kernel(const double* __restrict__ W, const double* __restrict__ G,
                                                       const double* __restrict__ Z0, const double* __restrict__ U,
                                                       long long u_stride_t, long long u_stride_b, int n, int r, int d,
                                                       int nb, double* __restrict__ out) {
    const int t = blockIdx.x + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const double* Wt = W + (size_t)t * n * d;
    for (int o = warp; o < nb * n; o += 4) {
        const int b = o / n, i = o - b * n;
        double s = 0.0;
        for (int q = lane; q < d; q += 32) s = fma(Wt[(size_t)i * d + q], Z0[(size_t)b * d + q], s);
        for (int e = lane; e < t * r; e += 32) {
            const int sidx = e / r, c = e - sidx * r;
            s = fma(G[((size_t)(t - 1 - sidx) * n + i) * r + c], U[sidx * u_stride_t + b * u_stride_b + c], s);
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_down_sync(0xffffffffu, s, off);
        if (lane == 0) out[((size_t)(t - 1) * nb + b) * n + i] = s;
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------------
struct brov_koopman {
    int device, n, r, k, d, num_sms;
    double gamma;
    double *C, *c2, *A, *B;     // device copies
    double* W;                  // [cap_t + 1][n][d]: W_0 .. W_{have_t}
    double* G;                  // [cap_t][n][r]:     G_0 .. G_{have_t - 1}
    int cap_t, have_t;
    double* partial;
    size_t cap_partial;
    double* Z0;
    size_t cap_z0;
    double* P;                  // [n_horizons][windows][n] lifted part of the predictions (multi-horizon scoring)
    size_t cap_P;
    int* d_Hs;                  // [BROV_MAX_H]
};

static int koop_prepare(brov_koopman* h, int T, cudaStream_t st) {
    const int n = h->n, d = h->d, r = h->r;
    if (T > h->cap_t) {
        int cap = h->cap_t ? h->cap_t : 128;
        while (cap < T) cap *= 2;
        double *W = nullptr, *G = nullptr;
        BROV_CUDA_TRY(cudaMalloc(&W, (size_t)(cap + 1) * n * d * sizeof(double)));
        if (cudaMalloc(&G, (size_t)cap * n * r * sizeof(double)) != cudaSuccess) {
            cudaFree(W);
            return brov::fail_msg(BROV_ENOMEM, "out of device memory for %d Koopman gain blocks", cap);
        }
        if (h->W) {
            BROV_CUDA_TRY(cudaMemcpyAsync(W, h->W, (size_t)(h->have_t + 1) * n * d * sizeof(double), cudaMemcpyDeviceToDevice, st));
            BROV_CUDA_TRY(cudaMemcpyAsync(G, h->G, (size_t)h->have_t * n * r * sizeof(double), cudaMemcpyDeviceToDevice, st));
            BROV_CUDA_TRY(cudaStreamSynchronize(st));
            cudaFree(h->W);
            cudaFree(h->G);
        } else {
            // W_0 = E = [I_n 0]
            std::vector<double> E((size_t)n * d, 0.0);
            for (int i = 0; i < n; ++i) E[(size_t)i * d + i] = 1.0;
            BROV_CUDA_TRY(cudaMemcpyAsync(W, E.data(), E.size() * sizeof(double), cudaMemcpyHostToDevice, st));
            BROV_CUDA_TRY(cudaStreamSynchronize(st));
        }
        h->W = W; h->G = G; h->cap_t = cap;
    }
    if (T <= h->have_t) return BROV_OK;
    const int j0 = h->have_t;
    const size_t pw_smem = ((size_t)d * NMAX + (size_t)PW_SLICES * NMAX * (PW_COLS + 1)) * sizeof(double);
    if (pw_smem > 200 * 1024) return brov::fail_msg(BROV_EUNSUPPORTED, "model with d = %d does not fit shared memory", d);
    if (pw_smem > 48 * 1024) BROV_CUDA_TRY(cudaFuncSetAttribute(koop_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw_smem));
    for (int t = j0; t < T; ++t)
        koop_power_kernel<<<(d + PW_COLS - 1) / PW_COLS, PW_COLS * PW_SLICES, pw_smem, st>>>(h->W + (size_t)t * n * d, h->A, h->W + (size_t)(t + 1) * n * d, n, d);
    koop_gain_kernel<<<T - j0, 128, 0, st>>>(h->W, h->B, h->G, n, r, d, j0);
    BROV_CUDA_TRY(cudaGetLastError());
    h->have_t = T;
    return BROV_OK;
}

extern "C" int brov_koopman_create(int device, int n, int r, int k, double gamma, const double* centers,
                                   const double* A, const double* B, brov_koopman_t** out) {
    if (!out) return brov::fail_msg(BROV_EINVAL, "out is NULL");
    *out = nullptr;
    if (!centers || !A || !B) return brov::fail_msg(BROV_EINVAL, "NULL argument");
    if (n < 1 || n > NMAX || r < 1 || r > RMAX || k < 0) return brov::fail_msg(BROV_EINVAL, "unsupported dimensions n=%d r=%d k=%d (n <= %d, r <= %d)", n, r, k, NMAX, RMAX);
    int ndev = 0;
    BROV_CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return brov::fail_msg(BROV_EINVAL, "device %d out of range (%d visible)", device, ndev);
    cudaDeviceProp prop;
    BROV_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return brov::fail_msg(BROV_EUNSUPPORTED, "libbrov is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
    BROV_CUDA_TRY(cudaSetDevice(device));
    brov_koopman* h = new (std::nothrow) brov_koopman();
    if (!h) return brov::fail_msg(BROV_ENOMEM, "out of host memory");
    memset(h, 0, sizeof(*h));
    h->device = device; h->n = n; h->r = r; h->k = k; h->d = n + k; h->gamma = gamma;
    h->num_sms = prop.multiProcessorCount;
    const int d = h->d;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(&h->C, (size_t)(k > 0 ? k : 1) * n * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->c2, (size_t)(k > 0 ? k : 1) * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->A, (size_t)d * d * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc(&h->B, (size_t)d * r * sizeof(double));
    if (e == cudaSuccess && k > 0) e = cudaMemcpy(h->C, centers, (size_t)k * n * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->A, A, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(h->B, B, (size_t)d * r * sizeof(double), cudaMemcpyHostToDevice);
    if (e == cudaSuccess && k > 0) {
        koop_c2_kernel<<<(k + 127) / 128, 128>>>(h->C, h->c2, n, k);
        e = cudaDeviceSynchronize();
    }
    if (e != cudaSuccess) {
        cudaFree(h->C); cudaFree(h->c2); cudaFree(h->A); cudaFree(h->B);
        delete h;
        return brov::fail_msg(BROV_ECUDA, "brov_koopman_create: %s", cudaGetErrorString(e));
    }
    *out = h;
    return BROV_OK;
}

extern "C" void brov_koopman_destroy(brov_koopman_t* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaFree(h->C); cudaFree(h->c2); cudaFree(h->A); cudaFree(h->B); cudaFree(h->W); cudaFree(h->G);
    cudaFree(h->partial); cudaFree(h->Z0); cudaFree(h->P); cudaFree(h->d_Hs);
    delete h;
}

extern "C" int brov_koopman_lift(brov_koopman_t* h, const double* X_dev, long long rows, double* Z_dev, void* stream) {
    if (!h || (rows > 0 && (!X_dev || !Z_dev))) return brov::fail_msg(BROV_EINVAL, "NULL argument");
    if (rows <= 0) return BROV_OK;
    BROV_CUDA_TRY(cudaSetDevice(h->device));
    const size_t smem = ((size_t)h->k * h->n + h->k) * sizeof(double);
    if (smem > 200 * 1024) return brov::fail_msg(BROV_EUNSUPPORTED, "%d centers do not fit shared memory", h->k);
    if (smem > 48 * 1024) BROV_CUDA_TRY(cudaFuncSetAttribute(koop_lift_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    koop_lift_kernel<<<(unsigned)((rows + KB - 1) / KB), KB, smem, (cudaStream_t)stream>>>(X_dev, h->C, h->c2, h->gamma, rows, h->n, h->k, Z_dev);
    BROV_CUDA_TRY(cudaGetLastError());
    return BROV_OK;
}

template <int N, int NH>
static int koop_liftw_launch(brov_koopman* h, const double* X, long long nwin, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024) BROV_CUDA_TRY(cudaFuncSetAttribute(koop_liftw_kernel<N, NH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long ntiles = (nwin + LWT * LWW - 1) / (LWT * LWW);
    const unsigned grid = (unsigned)(ntiles < h->num_sms ? ntiles : h->num_sms);
    koop_liftw_kernel<N, NH><<<grid, LWT, smem, st>>>(X, h->C, h->c2, h->W, h->d_Hs, h->gamma, nwin, h->k, h->P);
    BROV_CUDA_TRY(cudaGetLastError());
    return BROV_OK;
}
template <int N>
static int koop_liftw_nh(brov_koopman* h, int nh, const double* X, long long nwin, size_t smem, cudaStream_t st) {
    switch (nh) {
        case 1: return koop_liftw_launch<N, 1>(h, X, nwin, smem, st);
        case 2: return koop_liftw_launch<N, 2>(h, X, nwin, smem, st);
        case 3: return koop_liftw_launch<N, 3>(h, X, nwin, smem, st);
    }
    return brov::fail_msg(BROV_EINVAL, "internal: %d horizons per lift pass", nh);
}
template <int N, int R>
static int koop_fir_launch(brov_koopman* h, const double* X, const double* U, const double* P, int H, long long nwin,
                           long long rows, double* partial, cudaStream_t st) {
    const size_t g_bytes = (size_t)H * N * R * sizeof(double);
    const size_t u_bytes = (size_t)(FT * FW + H) * R * sizeof(double);
    const size_t budget = 220 * 1024;
    const int u_in_smem = u_bytes <= budget;                       // the input tile first: it removes the L1 bottleneck
    const int g_in_smem = g_bytes + (u_in_smem ? u_bytes : 0) <= budget;
    const size_t smem = (g_in_smem ? g_bytes : 0) + (u_in_smem ? u_bytes : 0);
    if (smem > 48 * 1024) BROV_CUDA_TRY(cudaFuncSetAttribute(koop_fir_se_kernel<N, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((nwin + FT * FW - 1) / (FT * FW));
    koop_fir_se_kernel<N, R><<<grid, FT, smem, st>>>(X, U, P, h->G, H, nwin, rows, g_in_smem, u_in_smem, partial);
    BROV_CUDA_TRY(cudaGetLastError());
    return BROV_OK;
}

// windows scored for horizon H: k in [0, min(win_limit, rows - H)); win_limit < 0 = no limit
static int koop_score(brov_koopman* h, const double* X_dev, const double* U_dev, long long rows, long long win_limit,
                      int n_horizons, const int* horizons, double* se_out_dev, void* stream) {
    if (!h || !X_dev || !U_dev || !se_out_dev || !horizons) return brov::fail_msg(BROV_EINVAL, "NULL argument");
    if (n_horizons < 1 || n_horizons > BROV_MAX_H) return brov::fail_msg(BROV_EINVAL, "n_horizons must be 1..%d", BROV_MAX_H);
    for (int q = 0; q < n_horizons; ++q)
        if (horizons[q] < 1 || (q && horizons[q] <= horizons[q - 1])) return brov::fail_msg(BROV_EINVAL, "horizons must be >= 1 and strictly ascending");
    if (!((h->n == 12 && (h->r == 8 || h->r == 6)) || (h->n == 13 && h->r == 6)))
        return brov::fail_msg(BROV_EUNSUPPORTED, "compiled for (n, r) = (12, 8), (12, 6), (13, 6); got (%d, %d)", h->n, h->r);
    cudaStream_t st = (cudaStream_t)stream;
    BROV_CUDA_TRY(cudaSetDevice(h->device));
    BROV_CUDA_TRY(cudaMemsetAsync(se_out_dev, 0, n_horizons * sizeof(double), st));
    long long nwin = rows - horizons[0];              // windows of the shortest horizon: the lift covers all of them
    if (win_limit >= 0 && nwin > win_limit) nwin = win_limit;
    if (nwin <= 0) return BROV_OK;
    int rc = koop_prepare(h, horizons[n_horizons - 1], st);
    if (rc) return rc;
    const int n = h->n, d = h->d;
    // how many horizons' decoder rows fit shared memory beside the centres
    const size_t base = ((size_t)h->k * n + h->k) * sizeof(double), per = (size_t)d * n * sizeof(double);
    int fit = (int)((220 * 1024 - (long long)base) / (long long)per);
    if (fit < 1) return brov::fail_msg(BROV_EUNSUPPORTED, "model with d = %d does not fit shared memory", d);
    if (fit > 3) fit = 3;
    const size_t needP = (size_t)fit * nwin * n;
    if (needP > h->cap_P) {
        cudaFree(h->P);
        h->P = nullptr; h->cap_P = 0;
        BROV_CUDA_TRY(cudaMalloc(&h->P, needP * sizeof(double)));
        h->cap_P = needP;
    }
    if (!h->d_Hs) BROV_CUDA_TRY(cudaMalloc(&h->d_Hs, BROV_MAX_H * sizeof(int)));
    const size_t nblocks = (size_t)((nwin + FT * FW - 1) / (FT * FW));
    if (nblocks > h->cap_partial) {
        cudaFree(h->partial);
        h->partial = nullptr; h->cap_partial = 0;
        BROV_CUDA_TRY(cudaMalloc(&h->partial, nblocks * sizeof(double)));
        h->cap_partial = nblocks;
    }
    for (int q0 = 0; q0 < n_horizons; q0 += fit) {
        const int nh = (n_horizons - q0 < fit) ? (n_horizons - q0) : fit;
        BROV_CUDA_TRY(cudaMemcpyAsync(h->d_Hs, horizons + q0, nh * sizeof(int), cudaMemcpyHostToDevice, st));
        const size_t smem = base + nh * per;
        rc = (n == 12) ? koop_liftw_nh<12>(h, nh, X_dev, nwin, smem, st) : koop_liftw_nh<13>(h, nh, X_dev, nwin, smem, st);
        if (rc) return rc;
        for (int q = 0; q < nh; ++q) {
            const int H = horizons[q0 + q];
            long long nw = rows - H;
            if (win_limit >= 0 && nw > win_limit) nw = win_limit;
            if (nw <= 0) continue;
            const double* Pq = h->P + (size_t)q * nwin * n;
            if (n == 12 && h->r == 8) rc = koop_fir_launch<12, 8>(h, X_dev, U_dev, Pq, H, nw, rows, h->partial, st);
            else if (n == 12) rc = koop_fir_launch<12, 6>(h, X_dev, U_dev, Pq, H, nw, rows, h->partial, st);
            else rc = koop_fir_launch<13, 6>(h, X_dev, U_dev, Pq, H, nw, rows, h->partial, st);
            if (rc) return rc;
            koop_finish_kernel<<<1, 256, 0, st>>>(h->partial, (int)((nw + FT * FW - 1) / (FT * FW)), se_out_dev + q0 + q);
            BROV_CUDA_TRY(cudaGetLastError());
        }
        if (q0 + fit < n_horizons) BROV_CUDA_TRY(cudaStreamSynchronize(st));   // d_Hs is rewritten by the next pass
    }
    return BROV_OK;
}

extern "C" int brov_koopman_multistep_se_multi(brov_koopman_t* h, const double* X_dev, const double* U_dev,
                                               long long rows, int n_horizons, const int* horizons,
                                               double* se_out_dev, void* stream) {
    return koop_score(h, X_dev, U_dev, rows, -1, n_horizons, horizons, se_out_dev, stream);
}

extern "C" int brov_koopman_multistep_se(brov_koopman_t* h, const double* X_dev, const double* U_dev, long long rows,
                                         long long n_windows, int H, double* se_out_dev, void* stream) {
    if (H < 1) return brov::fail_msg(BROV_EINVAL, "H must be >= 1");
    if (n_windows < 0 || n_windows + H > rows) return brov::fail_msg(BROV_EINVAL, "n_windows + H = %lld exceeds rows = %lld", n_windows + H, rows);
    return koop_score(h, X_dev, U_dev, rows, n_windows, 1, &H, se_out_dev, stream);
}

extern "C" int brov_koopman_simulate(brov_koopman_t* h, const double* X0_dev, const double* U_dev, long long T,
                                     long long nb, int u_shared, double* out_dev, void* stream) {
    if (!h || !X0_dev || !out_dev || (T > 0 && !U_dev)) return brov::fail_msg(BROV_EINVAL, "NULL argument");
    if (T <= 0 || nb <= 0) return BROV_OK;
    if (T > (1 << 20) || nb > (1 << 20)) return brov::fail_msg(BROV_EINVAL, "T or batch too large");
    cudaStream_t st = (cudaStream_t)stream;
    BROV_CUDA_TRY(cudaSetDevice(h->device));
    int rc = koop_prepare(h, (int)T, st);
    if (rc) return rc;
    const size_t need = (size_t)nb * h->d;
    if (need > h->cap_z0) {
        cudaFree(h->Z0);
        h->Z0 = nullptr; h->cap_z0 = 0;
        BROV_CUDA_TRY(cudaMalloc(&h->Z0, need * sizeof(double)));
        h->cap_z0 = need;
    }
    rc = brov_koopman_lift(h, X0_dev, nb, h->Z0, stream);
    if (rc) return rc;
    const long long st_t = u_shared ? h->r : nb * h->r, st_b = u_shared ? 0 : h->r;
    koop_sim_kernel<<<(unsigned)T, 128, 0, st>>>(h->W, h->G, h->Z0, U_dev, st_t, st_b, h->n, h->r, h->d, (int)nb, out_dev);
    BROV_CUDA_TRY(cudaGetLastError());
    return BROV_OK;
}
