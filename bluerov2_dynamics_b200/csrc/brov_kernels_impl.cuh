// brov_kernels_impl.cuh — runtime -> template dispatch.  Included by brov_kernels_f32.cu / brov_kernels_f64.cu with
// BROV_SCALAR defined; each instantiates every (model, integrator, lag1, per-vehicle) combination for its type.
#pragma once
#include "brov_kernels.cuh"

namespace brov {

template <typename T, int MODEL, int INTEG, bool LAG1>
static cudaError_t rollout_go(const RolloutArgs<T>& a, cudaStream_t st) {
    constexpr int NX = ModelDim<MODEL>::NX;
    const int grid = (a.n + ROLLOUT_BLOCK - 1) / ROLLOUT_BLOCK;
    size_t smem = a.traj ? (size_t)(ROLLOUT_BLOCK / 32) * 32 * NX * sizeof(T) : 0;
    if (a.pv) {
        smem += (size_t)KP_COUNT * ROLLOUT_BLOCK * sizeof(T);
        auto kern = rollout_kernel<T, MODEL, INTEG, LAG1, true>;
        if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        kern<<<grid, ROLLOUT_BLOCK, smem, st>>>(a);
    } else {
        rollout_kernel<T, MODEL, INTEG, LAG1, false><<<grid, ROLLOUT_BLOCK, smem, st>>>(a);
    }
    return cudaGetLastError();
}

template <typename T, int MODEL, bool LAG1>
static cudaError_t rollout_integ(int integ, const RolloutArgs<T>& a, cudaStream_t st) {
    return integ == INTEG_RK4 ? rollout_go<T, MODEL, INTEG_RK4, LAG1>(a, st)
                              : rollout_go<T, MODEL, INTEG_EULER, LAG1>(a, st);
}

template <typename T>
cudaError_t launch_rollout(int model, int integ, bool lag1, const RolloutArgs<T>& a, cudaStream_t st) {
    if (a.n <= 0 || a.steps <= 0) return cudaSuccess;
    switch (model) {
        case MODEL_THRUSTER8: return rollout_integ<T, MODEL_THRUSTER8, false>(integ, a, st);
        case MODEL_WRENCH12:
            return lag1 ? rollout_integ<T, MODEL_WRENCH12, true>(integ, a, st)
                        : rollout_integ<T, MODEL_WRENCH12, false>(integ, a, st);
        case MODEL_QUAT13:
            return lag1 ? rollout_integ<T, MODEL_QUAT13, true>(integ, a, st)
                        : rollout_integ<T, MODEL_QUAT13, false>(integ, a, st);
    }
    return cudaErrorInvalidValue;
}

template <typename T, int MODEL, bool LAG1>
static cudaError_t rhs_go(const RhsArgs<T>& a, cudaStream_t st) {
    const int grid = (a.n + ROLLOUT_BLOCK - 1) / ROLLOUT_BLOCK;
    if (a.pv) {
        size_t smem = (size_t)KP_COUNT * ROLLOUT_BLOCK * sizeof(T);
        rhs_kernel<T, MODEL, LAG1, true><<<grid, ROLLOUT_BLOCK, smem, st>>>(a);
    } else {
        rhs_kernel<T, MODEL, LAG1, false><<<grid, ROLLOUT_BLOCK, 0, st>>>(a);
    }
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_rhs(int model, bool lag1, const RhsArgs<T>& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    switch (model) {
        case MODEL_THRUSTER8: return rhs_go<T, MODEL_THRUSTER8, false>(a, st);
        case MODEL_WRENCH12: return lag1 ? rhs_go<T, MODEL_WRENCH12, true>(a, st) : rhs_go<T, MODEL_WRENCH12, false>(a, st);
        case MODEL_QUAT13: return lag1 ? rhs_go<T, MODEL_QUAT13, true>(a, st) : rhs_go<T, MODEL_QUAT13, false>(a, st);
    }
    return cudaErrorInvalidValue;
}

template <typename T, int MODEL>
static cudaError_t se_go(int integ, const SeArgs<T>& a, int nblocks, cudaStream_t st) {
    if (integ == INTEG_RK4) se_kernel<T, MODEL, INTEG_RK4><<<nblocks, SE_BLOCK, 0, st>>>(a);
    else se_kernel<T, MODEL, INTEG_EULER><<<nblocks, SE_BLOCK, 0, st>>>(a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_se(int model, int integ, const SeArgs<T>& a, int nblocks, double* se_out, cudaStream_t st) {
    cudaError_t e = cudaErrorInvalidValue;
    switch (model) {
        case MODEL_THRUSTER8: e = se_go<T, MODEL_THRUSTER8>(integ, a, nblocks, st); break;
        case MODEL_WRENCH12: e = se_go<T, MODEL_WRENCH12>(integ, a, nblocks, st); break;
        case MODEL_QUAT13: e = se_go<T, MODEL_QUAT13>(integ, a, nblocks, st); break;
    }
    if (e != cudaSuccess) return e;
    se_finish_kernel<<<1, 256, 0, st>>>(a.partial, nblocks, se_out);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_reduced9(const Red9Consts<T>& c, const T* X, const T* U, T* O, long long B, cudaStream_t st) {
    if (B <= 0) return cudaSuccess;
    const long long grid = (B + RED9_BLOCK - 1) / RED9_BLOCK;
    const int vec = ((reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(O)) & 15) == 0;
    reduced9_kernel<T><<<(unsigned)grid, RED9_BLOCK, 0, st>>>(c, X, U, O, B, vec);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_fma_peak(int iters, int blocks, T* scratch, cudaStream_t st) {
    fma_peak_kernel<T><<<blocks, 256, 0, st>>>(scratch, iters, T(0.999), T(1e-3));
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_thruster_wrench(const ThrusterArgs<T>& a, cudaStream_t st) {
    if (a.n <= 0) return cudaSuccess;
    thruster_wrench_kernel<T><<<(a.n + ROLLOUT_BLOCK - 1) / ROLLOUT_BLOCK, ROLLOUT_BLOCK, 0, st>>>(a);
    return cudaGetLastError();
}

#define BROV_INSTANTIATE(T)                                                                                         \
    template cudaError_t launch_rollout<T>(int, int, bool, const RolloutArgs<T>&, cudaStream_t);                   \
    template cudaError_t launch_rhs<T>(int, bool, const RhsArgs<T>&, cudaStream_t);                                \
    template cudaError_t launch_se<T>(int, int, const SeArgs<T>&, int, double*, cudaStream_t);                     \
    template cudaError_t launch_reduced9<T>(const Red9Consts<T>&, const T*, const T*, T*, long long, cudaStream_t); \
    template cudaError_t launch_fma_peak<T>(int, int, T*, cudaStream_t);                                           \
    template cudaError_t launch_thruster_wrench<T>(const ThrusterArgs<T>&, cudaStream_t);

}  // namespace brov
