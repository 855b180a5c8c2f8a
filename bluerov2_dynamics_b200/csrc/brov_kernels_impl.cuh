// brov_kernels_impl.cuh — runtime -> template dispatch.  Included by brov_kernels_f32.cu / brov_kernels_f64.cu; each
// instantiates every (model, integrator, lag1, per-vehicle, lag representation) combination for its scalar type.
#pragma once
#include "brov_kernels.cuh"

namespace brov {

template <typename T, int MODEL, int INTEG, bool LAG1, bool LAGW>
static cudaError_t rollout_go(const RolloutArgs<T>& a, cudaStream_t st) {
    constexpr int BLOCK = BlockOf<T>::N;
    constexpr int NX = ModelDim<MODEL>::NX;
    using LR = LagRegs<T, MODEL, LAG1, LAGW>;
    const int grid = ((a.n + BLOCK - 1) / BLOCK) * (a.quanta > 1 ? a.quanta : 1);
    size_t smem = a.traj ? (size_t)(BLOCK / 32) * 32 * NX * sizeof(T) : 0;
    if (LR::SMEM) smem += (size_t)LR::N * BLOCK * sizeof(T);
    if (AccInSmem<T>::V && INTEG == INTEG_RK4) smem += (size_t)NX * BLOCK * sizeof(T);
    if (a.pv) {
        smem += (size_t)KP_COUNT * BLOCK * sizeof(T);
        auto kern = rollout_kernel<T, MODEL, INTEG, LAG1, true, LAGW>;
        // static + dynamic shared memory beyond 48 KB needs the opt-in (the fp64 coefficient table alone is 36 KB)
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return e;
        if (smem + fa.sharedSizeBytes > 48 * 1024) {
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kern<<<grid, BLOCK, smem, st>>>(a);
    } else {
        rollout_kernel<T, MODEL, INTEG, LAG1, false, LAGW><<<grid, BLOCK, smem, st>>>(a);
    }
    return cudaGetLastError();
}

// resident blocks per SM of the kernel that launch_rollout would pick (for the temporal-tiling heuristic)
template <typename T, int MODEL, int INTEG, bool LAG1, bool LAGW>
static int rollout_occ(bool pv, bool traj) {
    constexpr int BLOCK = BlockOf<T>::N;
    constexpr int NX = ModelDim<MODEL>::NX;
    using LR = LagRegs<T, MODEL, LAG1, LAGW>;
    size_t smem = traj ? (size_t)(BLOCK / 32) * 32 * NX * sizeof(T) : 0;
    if (LR::SMEM) smem += (size_t)LR::N * BLOCK * sizeof(T);
    if (AccInSmem<T>::V && INTEG == INTEG_RK4) smem += (size_t)NX * BLOCK * sizeof(T);
    if (pv) smem += (size_t)KP_COUNT * BLOCK * sizeof(T);
    int nb = 0;
    cudaError_t e = pv ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rollout_kernel<T, MODEL, INTEG, LAG1, true, LAGW>, BLOCK, smem)
                       : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, rollout_kernel<T, MODEL, INTEG, LAG1, false, LAGW>, BLOCK, smem);
    return e == cudaSuccess ? nb : 0;
}

template <typename T>
int rollout_blocks_per_sm(int model, int integ, bool lag1, bool lagw, bool pv, bool traj) {
    const bool rk4 = integ == INTEG_RK4;
    switch (model) {
        case MODEL_THRUSTER8:
            if (lagw) return rk4 ? rollout_occ<T, MODEL_THRUSTER8, INTEG_RK4, false, true>(pv, traj) : rollout_occ<T, MODEL_THRUSTER8, INTEG_EULER, false, true>(pv, traj);
            return rk4 ? rollout_occ<T, MODEL_THRUSTER8, INTEG_RK4, false, false>(pv, traj) : rollout_occ<T, MODEL_THRUSTER8, INTEG_EULER, false, false>(pv, traj);
        case MODEL_WRENCH12:
            if (lag1) return rk4 ? rollout_occ<T, MODEL_WRENCH12, INTEG_RK4, true, false>(pv, traj) : rollout_occ<T, MODEL_WRENCH12, INTEG_EULER, true, false>(pv, traj);
            return rk4 ? rollout_occ<T, MODEL_WRENCH12, INTEG_RK4, false, false>(pv, traj) : rollout_occ<T, MODEL_WRENCH12, INTEG_EULER, false, false>(pv, traj);
        case MODEL_QUAT13:
            if (lag1) return rk4 ? rollout_occ<T, MODEL_QUAT13, INTEG_RK4, true, false>(pv, traj) : rollout_occ<T, MODEL_QUAT13, INTEG_EULER, true, false>(pv, traj);
            return rk4 ? rollout_occ<T, MODEL_QUAT13, INTEG_RK4, false, false>(pv, traj) : rollout_occ<T, MODEL_QUAT13, INTEG_EULER, false, false>(pv, traj);
        case MODEL_DI12_U8:
            return rk4 ? rollout_occ<T, MODEL_DI12_U8, INTEG_RK4, false, false>(pv, traj) : rollout_occ<T, MODEL_DI12_U8, INTEG_EULER, false, false>(pv, traj);
        case MODEL_DI12_U6:
            return rk4 ? rollout_occ<T, MODEL_DI12_U6, INTEG_RK4, false, false>(pv, traj) : rollout_occ<T, MODEL_DI12_U6, INTEG_EULER, false, false>(pv, traj);
        case MODEL_DIQ13_U6:
            return rk4 ? rollout_occ<T, MODEL_DIQ13_U6, INTEG_RK4, false, false>(pv, traj) : rollout_occ<T, MODEL_DIQ13_U6, INTEG_EULER, false, false>(pv, traj);
    }
    return 0;
}
template <typename T> int rollout_block_threads() { return BlockOf<T>::N; }

template <typename T, int MODEL, bool LAG1, bool LAGW>
static cudaError_t rollout_integ(int integ, const RolloutArgs<T>& a, cudaStream_t st) {
    return integ == INTEG_RK4 ? rollout_go<T, MODEL, INTEG_RK4, LAG1, LAGW>(a, st)
                              : rollout_go<T, MODEL, INTEG_EULER, LAG1, LAGW>(a, st);
}

template <typename T>
cudaError_t launch_rollout(int model, int integ, bool lag1, bool lagw, const RolloutArgs<T>& a, cudaStream_t st) {
    if (a.n <= 0 || a.steps <= 0) return cudaSuccess;
    switch (model) {
        case MODEL_THRUSTER8:
            return lagw ? rollout_integ<T, MODEL_THRUSTER8, false, true>(integ, a, st)
                        : rollout_integ<T, MODEL_THRUSTER8, false, false>(integ, a, st);
        case MODEL_WRENCH12:
            return lag1 ? rollout_integ<T, MODEL_WRENCH12, true, false>(integ, a, st)
                        : rollout_integ<T, MODEL_WRENCH12, false, false>(integ, a, st);
        case MODEL_QUAT13:
            return lag1 ? rollout_integ<T, MODEL_QUAT13, true, false>(integ, a, st)
                        : rollout_integ<T, MODEL_QUAT13, false, false>(integ, a, st);
        case MODEL_DI12_U8: return rollout_integ<T, MODEL_DI12_U8, false, false>(integ, a, st);
        case MODEL_DI12_U6: return rollout_integ<T, MODEL_DI12_U6, false, false>(integ, a, st);
        case MODEL_DIQ13_U6: return rollout_integ<T, MODEL_DIQ13_U6, false, false>(integ, a, st);
    }
    return cudaErrorInvalidValue;
}

template <typename T, int MODEL, bool LAG1>
static cudaError_t rhs_go(const RhsArgs<T>& a, cudaStream_t st) {
    const int grid = (a.n + RHS_BLOCK - 1) / RHS_BLOCK;
    if (a.pv) {
        size_t smem = (size_t)KP_COUNT * RHS_BLOCK * sizeof(T);
        rhs_kernel<T, MODEL, LAG1, true><<<grid, RHS_BLOCK, smem, st>>>(a);
    } else {
        rhs_kernel<T, MODEL, LAG1, false><<<grid, RHS_BLOCK, 0, st>>>(a);
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
        case MODEL_DI12_U8: return rhs_go<T, MODEL_DI12_U8, false>(a, st);
        case MODEL_DI12_U6: return rhs_go<T, MODEL_DI12_U6, false>(a, st);
        case MODEL_DIQ13_U6: return rhs_go<T, MODEL_DIQ13_U6, false>(a, st);
    }
    return cudaErrorInvalidValue;
}

template <typename T> int se_blocks(long long nwin) {
    return (int)((nwin + BlockOf<T>::N - 1) / BlockOf<T>::N);
}

template <typename T, int MODEL>
static cudaError_t se_go(int integ, const SeArgs<T>& a, cudaStream_t st) {
    const int nblocks = se_blocks<T>(a.nwin);
    if (integ == INTEG_RK4) se_kernel<T, MODEL, INTEG_RK4><<<nblocks, BlockOf<T>::N, 0, st>>>(a);
    else se_kernel<T, MODEL, INTEG_EULER><<<nblocks, BlockOf<T>::N, 0, st>>>(a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_se(int model, int integ, const SeArgs<T>& a, double* se_out, cudaStream_t st) {
    cudaError_t e = cudaErrorInvalidValue;
    switch (model) {
        case MODEL_THRUSTER8: e = se_go<T, MODEL_THRUSTER8>(integ, a, st); break;
        case MODEL_WRENCH12: e = se_go<T, MODEL_WRENCH12>(integ, a, st); break;
        case MODEL_QUAT13: e = se_go<T, MODEL_QUAT13>(integ, a, st); break;
        case MODEL_DI12_U8: e = se_go<T, MODEL_DI12_U8>(integ, a, st); break;
        case MODEL_DI12_U6: e = se_go<T, MODEL_DI12_U6>(integ, a, st); break;
        case MODEL_DIQ13_U6: e = se_go<T, MODEL_DIQ13_U6>(integ, a, st); break;
    }
    if (e != cudaSuccess) return e;
    se_finish_kernel<<<1, 256, 0, st>>>(a.partial, se_blocks<T>(a.nwin), se_out);
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
    thruster_wrench_kernel<T><<<(a.n + RHS_BLOCK - 1) / RHS_BLOCK, RHS_BLOCK, 0, st>>>(a);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_thruster_series(const ThrusterSeriesArgs<T>& a, cudaStream_t st) {
    if (a.rows <= 0) return cudaSuccess;
    thruster_series_kernel<T><<<(unsigned)((a.rows + RHS_BLOCK - 1) / RHS_BLOCK), RHS_BLOCK, 0, st>>>(a);
    return cudaGetLastError();
}

#define BROV_INSTANTIATE(T)                                                                                         \
    template cudaError_t launch_rollout<T>(int, int, bool, bool, const RolloutArgs<T>&, cudaStream_t);             \
    template int rollout_blocks_per_sm<T>(int, int, bool, bool, bool, bool);                                      \
    template int rollout_block_threads<T>();                                                                       \
    template cudaError_t launch_rhs<T>(int, bool, const RhsArgs<T>&, cudaStream_t);                                \
    template cudaError_t launch_se<T>(int, int, const SeArgs<T>&, double*, cudaStream_t);                          \
    template int se_blocks<T>(long long);                                                                          \
    template cudaError_t launch_reduced9<T>(const Red9Consts<T>&, const T*, const T*, T*, long long, cudaStream_t); \
    template cudaError_t launch_fma_peak<T>(int, int, T*, cudaStream_t);                                           \
    template cudaError_t launch_thruster_wrench<T>(const ThrusterArgs<T>&, cudaStream_t);                          \
    template cudaError_t launch_thruster_series<T>(const ThrusterSeriesArgs<T>&, cudaStream_t);

}  // namespace brov
