// brov_pinc_tc.cuh — the dense layers of the PINc network on the 5th-generation tensor cores (tcgen05 + TMEM).
// Included by brov_pinc.cu inside its anonymous namespace (it reuses softplus_f, thruster_map4, x12_to_9, ...).
//
// The four 64 x 64 layers (plus 14 -> 64 and 64 -> 9) over a block of windows are small GEMMs with the SAME weights
// for every window and every step: one CTA scores a tile of 128 windows, thread t owns window t = accumulator row t.
//   D[128 x N] (TMEM, fp32) = A[128 x K] (activations, shared memory) * W[N x K]^T (shared memory, staged once)
// Operands are TF32.  A plain TF32 product (10-bit mantissas) would miss the 2e-6 forward tolerance against the
// reference's float32 torch network by three orders of magnitude, so every operand is split into two TF32 numbers,
// x = x_hi + x_lo with x_hi = rna_tf32(x), x_lo = rna_tf32(x - x_hi), and a layer is the three products
// A_hi W_hi + A_hi W_lo + A_lo W_hi accumulated in fp32 (the dropped A_lo W_lo term and the rounding of the lo parts
// are each below 2^-22 relative per product, unbiased) — 3 x K/8 tcgen05.mma instructions per layer, issued by one
// thread.  Between layers every thread pulls its 64 accumulators back with tcgen05.ld and does bias, softplus (two
// MUFU), LayerNorm and the hi/lo split in registers, then writes its row of the next A straight back into TENSOR
// MEMORY with tcgen05.st (the A operand of tcgen05.mma may live in TMEM: lane = row, column = k): activations never
// touch shared memory.  The weights are staged once per CTA in shared memory in the canonical K-major core-matrix
// layout the UMMA descriptors describe (8 rows x 16 bytes per core matrix, no swizzle).
//
// A CTA is TWO independent tiles (threads 0..127 and 128..255), each with its own mbarrier, named barrier and TMEM
// columns: while one tile's MMAs run (or wait to be issued), the other tile's epilogue keeps the CUDA cores and MUFU
// busy — a first version with one tile per CTA and A in shared memory (183 KB) serialised the two and was no faster
// than the CUDA-core kernel (73 against 79 ms on the 1M-window table).
// Shared memory per CTA: weights 115 KB (hi + lo of five layers) + parameters 3 KB, one CTA per SM, persistent over
// tiles.  TMEM per tile: 64 columns of accumulators + 64 + 64 columns of A_hi / A_lo = 192; 512 allocated per CTA.
#pragma once

constexpr int TC_M = 128;                       // windows per tile = threads per CTA = TMEM lanes
constexpr int TC_L0_HI = 0, TC_L0_LO = 1024;    // [64 n][16 k] each, core-matrix layout (K = 14 padded to 16)
constexpr int TC_L1 = 2048;                     // hidden layers l = 1..3: hi at TC_L1 + (l-1)*8192, lo 4096 later
constexpr int TC_L4_HI = TC_L1 + 3 * 8192, TC_L4_LO = TC_L4_HI + 1024;   // [16 n][64 k] each (9 outputs padded to 16)
constexpr int TC_PAR = TC_L4_LO + 1024;         // per hidden layer: bias[64], ln_w[64], ln_b[64]; then b4[16]
constexpr int TC_NW = TC_PAR + 4 * 192 + 16;    // floats in the blob
constexpr int TC_TILES = 2;                     // tiles (halves of 128 threads) per CTA
constexpr int TC_COLS = 192;                    // TMEM columns per tile: D [0,64), A_hi [64,128), A_lo [128,192)
constexpr size_t TC_SMEM_BYTES = (size_t)TC_NW * sizeof(float) + 64;

// float offset of element (row, k) of a K-major operand with `rows` rows in the canonical no-swizzle layout:
// core matrix = 8 rows x 4 floats (16 B per row, 128 B); core matrices of one K chunk follow each other along the
// rows (SBO = 128 B); the next K chunk starts rows/8 core matrices later (LBO = rows/8 * 128 B).
__host__ __device__ constexpr int tc_off(int rows, int row, int k) {
    return (k >> 2) * (rows >> 3) * 32 + (row >> 3) * 32 + (row & 7) * 4 + (k & 3);
}

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset >> 4 in [16,30) (K-chunk stride), stride byte offset >> 4 in [32,46) (8-row group stride),
// version 1 in [46,48), layout type 0 = no swizzle in [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor (InstrDescriptor): D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), both K-major,
// N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T: one instruction covers K = 8 (TF32)
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_bar_init(uint32_t bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_bar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity)
                     : "memory");
    } while (!ok);
}
// 16 consecutive fp32 accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ float tc_rna(float x) {   // round to TF32 (nearest, ties away), result in fp32 container
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// 16 consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
        "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
        "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
        "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
// 16 activations -> their TF32 hi / lo parts -> columns [k0, k0 + 16) of this thread's rows of A_hi / A_lo
__device__ __forceinline__ void tc_put16(uint32_t a_hi, uint32_t a_lo, int k0, const float* x) {
    float hi[16], lo[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        hi[j] = tc_rna(x[j]);
        lo[j] = tc_rna(x[j] - hi[j]);
    }
    tc_st16(a_hi + k0, hi);
    tc_st16(a_lo + k0, lo);
}

struct TcCtx {
    const float* sw;   // weights + parameters in shared memory (TC_NW floats)
    uint32_t bar;      // this tile's mbarrier (shared-window address): the MMA commits arrive on it
    uint32_t tm;       // this thread's TMEM address of the tile's column 0 (lane bits = first lane of its warp)
    uint32_t tm0;      // the tile's column 0 at lane 0 (what the MMA instructions address)
    uint32_t phase;    // parity the next wait expects
    int row;           // this thread's row of the tile = its window
    int half;          // tile of the CTA this thread belongs to
};

// One layer's products, issued by ONE thread after the tile's threads have synchronised on the freshly written A:
// D = A_hi W_hi^T + A_hi W_lo^T + A_lo W_hi^T over K (multiple of 8), N output columns.
__device__ __forceinline__ void tc_issue_layer(const TcCtx& c, int whi_off, int wlo_off, int K, int N) {
    const uint32_t b_lbo = (uint32_t)(N / 8) * 128;
    const uint32_t whi = tc_smem(c.sw + whi_off), wlo = tc_smem(c.sw + wlo_off);
    const uint32_t idesc = tc_idesc(TC_M, N);
    bool acc = false;
#pragma unroll 1
    for (int prod = 0; prod < 3; ++prod) {
        const uint32_t a0 = c.tm0 + (prod == 2 ? 128 : 64), b0 = prod == 1 ? wlo : whi;
        for (int k8 = 0; k8 < K / 8; ++k8) {     // A: 8 TMEM columns per instruction; B: two 16-byte K chunks
            tc_mma(c.tm0, a0 + 8 * k8, tc_desc(b0 + k8 * 2 * b_lbo, b_lbo, 128), idesc, acc);
            acc = true;
        }
    }
    tc_commit(c.bar);
}

// Runs one layer for the tile: every thread has stored its row of A.  Returns with the accumulators readable.
__device__ __forceinline__ void tc_layer(TcCtx& c, int whi_off, int wlo_off, int K, int N) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + c.half), "r"(TC_M) : "memory");     // the tile's 128 threads
    if (c.row == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc_issue_layer(c, whi_off, wlo_off, K, N);
    }
    tc_bar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// bias + AdaptiveSoftplus + LayerNorm on this thread's 64 accumulators (read back 16 at a time), result split and
// stored as this thread's row of the next layer's A
__device__ __forceinline__ void tc_hidden_epilogue(const TcCtx& c, int layer, float beta) {
    const float* par = c.sw + TC_PAR + layer * 192;
    const float ib = 1.0f / (beta + 1e-12f);
    float a[HID];
    float mean = 0.0f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float v[16];
        tc_ld16(c.tm + 16 * q, v);
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 b = *reinterpret_cast<const float4*>(par + 16 * q + j);
            a[16 * q + j + 0] = softplus_f(beta * (v[j + 0] + b.x)) * ib;
            a[16 * q + j + 1] = softplus_f(beta * (v[j + 1] + b.y)) * ib;
            a[16 * q + j + 2] = softplus_f(beta * (v[j + 2] + b.z)) * ib;
            a[16 * q + j + 3] = softplus_f(beta * (v[j + 3] + b.w)) * ib;
            mean += (a[16 * q + j] + a[16 * q + j + 1]) + (a[16 * q + j + 2] + a[16 * q + j + 3]);
        }
    }
    mean *= (1.0f / HID);
    float var = 0.0f;
#pragma unroll
    for (int j = 0; j < HID; ++j) {
        a[j] -= mean;
        var = fmaf(a[j], a[j], var);
    }
    const float rstd = rsqrtf(var * (1.0f / HID) + 1e-5f);
    // the tile's MMAs of this layer are complete (every thread waited on the commit): A may be overwritten
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float o[16];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
            const float4 g = *reinterpret_cast<const float4*>(par + 64 + 16 * q + j);
            const float4 b = *reinterpret_cast<const float4*>(par + 128 + 16 * q + j);
            o[j + 0] = fmaf(a[16 * q + j + 0] * rstd, g.x, b.x);
            o[j + 1] = fmaf(a[16 * q + j + 1] * rstd, g.y, b.y);
            o[j + 2] = fmaf(a[16 * q + j + 2] * rstd, g.z, b.z);
            o[j + 3] = fmaf(a[16 * q + j + 3] * rstd, g.w, b.w);
        }
        tc_put16(c.tm + 64, c.tm + 128, 16 * q, o);
    }
}

// PINcNet.forward (training/train_tank_brov2_rk4.py:627-673) for the tile's 128 windows: z[14] -> xn[9] per thread.
// Every thread of the tile must call it (tile-wide barriers inside).
__device__ __forceinline__ void pinc_forward_tc(TcCtx& c, const float* beta, const float (&z)[NIN], float (&xn)[9]) {
    {   // layer 0: A = z padded to K = 16
        float z16[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z16[j] = j < NIN ? z[j] : 0.0f;
        tc_put16(c.tm + 64, c.tm + 128, 0, z16);
    }
    tc_layer(c, TC_L0_HI, TC_L0_LO, 16, HID);
    tc_hidden_epilogue(c, 0, beta[0]);
#pragma unroll 1
    for (int l = 1; l <= 3; ++l) {
        tc_layer(c, TC_L1 + (l - 1) * 8192, TC_L1 + (l - 1) * 8192 + 4096, HID, HID);
        tc_hidden_epilogue(c, l, beta[l]);
    }
    tc_layer(c, TC_L4_HI, TC_L4_LO, HID, 16);
    float dx[16];
    tc_ld16(c.tm, dx);
    const float* b4 = c.sw + TC_PAR + 4 * 192;
#pragma unroll
    for (int j = 0; j < 9; ++j) dx[j] += b4[j];
    // residual update; body-frame (dx, dy) rotated by the CURRENT yaw; (cos, sin) re-normalised (:639-673)
    const float cs = z[3], sn = z[4];
    float base[9];
#pragma unroll
    for (int j = 0; j < 9; ++j) base[j] = z[j] + dx[j];
    xn[0] = (cs * dx[0] - sn * dx[1]) + z[0];
    xn[1] = (sn * dx[0] + cs * dx[1]) + z[1];
    xn[2] = base[2];
    const float nrm = fmaxf(sqrtf(base[3] * base[3] + base[4] * base[4]), 1e-6f);
    xn[3] = base[3] / nrm;
    xn[4] = base[4] / nrm;
#pragma unroll
    for (int j = 5; j < 9; ++j) xn[j] = base[j];
    // the next forward pass overwrites A and D: every thread has finished reading its accumulators (tcgen05.wait::ld)
    // before it arrives at the barrier inside the next tc_layer
}

// per-CTA setup / teardown (ALL threads of the CTA): stage the weights, init the tiles' mbarriers, allocate TMEM
// (warp 0 owns the allocation)
__device__ __forceinline__ void tc_setup(TcCtx& c, float* smem, const float* __restrict__ wtc, uint32_t* tmem_slot,
                                         uint64_t* bars) {
    c.sw = smem;
    for (int e = threadIdx.x * 4; e < TC_NW; e += blockDim.x * 4)
        *reinterpret_cast<float4*>(smem + e) = *reinterpret_cast<const float4*>(wtc + e);
    c.half = threadIdx.x / TC_M;
    c.row = threadIdx.x % TC_M;
    c.bar = tc_smem(bars + c.half);
    if (c.row == 0) tc_bar_init(c.bar);
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem(tmem_slot)), "r"(512u)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // the staged weights -> async proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    c.tm0 = *tmem_slot + (uint32_t)(c.half * TC_COLS);
    c.tm = c.tm0 + ((uint32_t)(c.row & ~31) << 16);                  // a warp addresses the 32 lanes of its quarter
    c.phase = 0;
}
__device__ __forceinline__ void tc_teardown(const TcCtx& c, const uint32_t* tmem_slot) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
}
