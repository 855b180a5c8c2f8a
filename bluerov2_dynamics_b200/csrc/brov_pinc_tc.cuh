// brov_pinc_tc.cuh — the dense layers of the PINc network on the 5th-generation tensor cores (tcgen05 + TMEM).
// Included by brov_pinc.cu inside its anonymous namespace (it reuses thruster_map4, x12_to_9, ...).
//
// The four 64 x 64 layers (plus 14 -> 64 and 64 -> 9) over a block of windows are small GEMMs with the SAME weights
// for every window and every step: a tile = 128 consecutive windows, thread t owns window t = accumulator row t =
// TMEM lane t.
//   D[128 x N] (TMEM, fp32) = A[128 x K] (activations, TMEM) * W[N x K]^T (shared memory, staged once per CTA)
// A plain TF32 product (10-bit mantissas) would miss the 2e-6 forward tolerance against the reference's float32 torch
// network by three orders of magnitude, so every operand is split, x = x_hi + x_lo with x_hi = rn_tf32(x), and a layer
// is three products accumulated in fp32:
//   D = A_hi(tf32) W_hi(tf32)^T + A_hi(tf32) W_lo(tf32)^T + A_lo(f16) W_hi(f16)^T        (A_lo W_lo dropped: 2^-22)
// The low part of an activation, |a_lo| <= 2^-11 |a|, needs no more than FP16's 11-bit significand, and its product
// runs as kind::f16 (K = 16 per instruction) into the same accumulator columns: 8 + 8 + 4 = 20 tcgen05.mma per hidden
// layer (M 128, N 64), issued by one thread per tile.
//
// Between layers every thread pulls its row's 64 accumulators back with tcgen05.ld, does bias, softplus (two MUFU),
// LayerNorm and the split in registers, and writes its row of the next A straight back into TENSOR MEMORY with
// tcgen05.st (the A operand of tcgen05.mma may live in TMEM: lane = row, column = k, two FP16 per column):
// activations never touch shared memory.  The weights sit in shared memory in the canonical K-major no-swizzle
// core-matrix layout the UMMA descriptors describe (8 rows x 16 bytes per core matrix).
//
// What limits the design is TENSOR MEMORY: a tile needs 64 (D) + 64 (A_hi) + 32 (A_lo as FP16) = 160 of the SM's 512
// columns, so a CTA runs THREE independent tiles (3 x 128 threads, each tile with its own mbarrier, named barrier and
// TMEM columns): while one tile's MMAs run, the other tiles' epilogues keep the CUDA cores and MUFU busy.  History on
// the 1M-window table (CUDA-core kernel: 79 ms): 1 tile x 128 threads, A in shared memory 73 ms; 2 tiles, A in TMEM
// 45 ms; 2 tiles x 2 threads per row 43 ms; folded scales + integer split 34.7 ms (ncu r02q: 29 % of the stall
// samples still waited for an MMA commit — two tiles do not cover each other's MMA -> epilogue chain; with TF32 low
// parts a tile needs 192 columns and a third does not fit); FP16 low parts, 3 tiles 27.3 ms.
//
// What the epilogue does NOT compute (folded on the host in double precision, brov_pinc_create): the LayerNorm affine
// (weight, bias) of layer l is multiplied into layer l + 1's weights and bias before the split; the activation's
// 1/beta and ln 2 factors are dropped because LayerNorm is scale-invariant up to its epsilon, which is rescaled
// instead; beta log2(e) multiplies the bias once.  AdaptiveSoftplus then is max(t, 0) + log2(1 + 2^-|t|) with
// t = fma(beta log2 e, acc, bias').  The TF32 rounding is two integer instructions (bits + 0x1000) & ~0x1fff — the
// PTX cvt.rna.tf32.f32 is emulated with four on sm_100a.
// Layer 0's K axis is permuted: k 0..8 = network state, 9 = dt, 10..11 = 0, 12..15 = the four projected thrust inputs
// (the host permutes W0 alike).
// Shared memory per CTA: TF32 weights 112 KB + FP16 copies of the high parts 28 KB + biases 1 KB = 145 KB, one CTA
// per SM, persistent over tiles.
#pragma once

constexpr int TC_M = 128;                       // windows per tile = threads per tile = TMEM lanes
constexpr int TC_TILES = 3;                     // tiles per CTA
constexpr int TC_THREADS = TC_TILES * TC_M;     // 384
constexpr int TC_COLS = 160;                    // TMEM columns per tile: D [0,64), A_hi [64,128), A_lo f16 [128,160)
// blob offsets in floats.  TF32 parts in [n][k] core-matrix layout:
constexpr int TC_L0_HI = 0, TC_L0_LO = 1024;    // [64 n][16 k] each (K = 14 padded to 16)
constexpr int TC_L1 = 2048;                     // hidden layers l = 1..3: hi at TC_L1 + (l-1)*8192, lo 4096 later
constexpr int TC_L4_HI = TC_L1 + 3 * 8192, TC_L4_LO = TC_L4_HI + 1024;   // [16 n][64 k] each (9 outputs padded to 16)
constexpr int TC_PAR = TC_L4_LO + 1024;         // per hidden layer: folded bias[64] (log2 domain); then b4'[16]
// FP16 copies of the weights' high parts (a float = two halves):
constexpr int TC_H16_L0 = TC_PAR + 4 * 64 + 16; // [64 n][16 k] halves
constexpr int TC_H16_L1 = TC_H16_L0 + 512;      // hidden layers l = 1..3: [64 n][64 k] halves, 2048 floats each
constexpr int TC_H16_L4 = TC_H16_L1 + 3 * 2048; // [16 n][64 k] halves
constexpr int TC_NW = TC_H16_L4 + 512;          // floats in the blob
constexpr size_t TC_SMEM_BYTES = (size_t)TC_NW * sizeof(float) + 64;
// per-layer scalars of the folded activation + LayerNorm (host: brov_pinc_create)
struct TcAct {
    float sc[4];    // beta * log2(e)
    float eps[4];   // 1e-5 * ((beta + 1e-12) / ln 2)^2: LayerNorm's epsilon in the un-scaled activation's units
    float sg[4];    // sign of ln 2 / (beta + 1e-12)
};
// input index (PINcNet: 9 state, 4 thrust, dt) held by column k of layer 0's A; -1 = zero padding
__host__ __device__ constexpr int tc_kmap0(int k) { return k < 9 ? k : (k == 9 ? 13 : (k < 12 ? -1 : k - 3)); }

// Offset of element (row, k) of a K-major operand with `rows` rows in the canonical no-swizzle layout.  A core matrix
// is 8 rows x 16 bytes (4 floats / 8 halves per row, 128 B); the core matrices of one K chunk follow each other along
// the rows (SBO = 128 B); the next K chunk starts rows/8 core matrices later (LBO = rows/8 * 128 B).
__host__ __device__ constexpr int tc_off(int rows, int row, int k) {          // in floats
    return (k >> 2) * (rows >> 3) * 32 + (row >> 3) * 32 + (row & 7) * 4 + (k & 3);
}
__host__ __device__ constexpr int tc_off16(int rows, int row, int k) {        // in halves
    return (k >> 3) * (rows >> 3) * 64 + (row >> 3) * 64 + (row & 7) * 8 + (k & 7);
}

__device__ __forceinline__ uint32_t tc_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// UMMA shared-memory descriptor (cute/arch/mma_sm100_desc.hpp SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset >> 4 in [16,30) (K-chunk stride), stride byte offset >> 4 in [32,46) (8-row group stride),
// version 1 in [46,48), layout type 0 = no swizzle in [61,64)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor (InstrDescriptor): D = F32 (1 << 4), A / B format in bits [7,10) / [10,13) (TF32 = 2,
// F16 = 0), both K-major, N >> 3 in bits [17,23), M >> 4 in bits [24,29)
__host__ __device__ constexpr uint32_t tc_idesc(int M, int N, bool tf32) {
    return (1u << 4) | (tf32 ? ((2u << 7) | (2u << 10)) : 0u) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]^T: one instruction covers K = 8 (kind::tf32) or K = 16 (kind::f16); either way 8
// TMEM columns of A and two K chunks of B
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc)
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
        // with a suspend-time hint the warp sleeps in hardware instead of polling (the polling loop was 7 % of the
        // issued instructions, r02v)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3; selp.u32 %0, 1, 0, p; }"
                     : "=r"(ok)
                     : "r"(bar), "r"(parity), "r"(4000u)
                     : "memory");
    } while (!ok);
}
// fp32 accumulator columns of this thread's TMEM lane.  Load and wait are ONE asm statement: nothing that consumes
// the registers can be scheduled between them.
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
}
// consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tc_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(
            taddr),
        "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]),
        "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
// two floats -> one register of two halves (round to nearest), `even` in the low half (the even K index)
__device__ __forceinline__ uint32_t tc_pack16(float even, float odd) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(odd), "f"(even));
    return r;
}

struct TcCtx {
    const float* sw;   // blob in shared memory (TC_NW floats)
    uint32_t bar;      // this tile's mbarrier (shared-window address): the MMA commits arrive on it
    uint32_t tm;       // this thread's TMEM address of the tile's column 0 (lane bits = first lane of its warp)
    uint32_t tm0;      // the tile's column 0 at lane 0 (what the MMA instructions address)
    uint32_t phase;    // parity the next wait expects
    int row;           // this thread's row of the tile = its window
    int tile;          // tile of the CTA this thread belongs to
};

// 16 activations -> A_hi columns [k0, k0 + 16) as TF32 numbers (nearest, ties away; finite inputs) and A_lo columns
// [k0 / 2, k0 / 2 + 8) as FP16 pairs of the exact remainders
__device__ __forceinline__ void tc_put16(const TcCtx& c, int k0, const float* x) {
    uint32_t hi[16], lo[8];
#pragma unroll
    for (int j = 0; j < 16; ++j) hi[j] = (__float_as_uint(x[j]) + 0x1000u) & 0xffffe000u;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        lo[j] = tc_pack16(x[2 * j] - __uint_as_float(hi[2 * j]), x[2 * j + 1] - __uint_as_float(hi[2 * j + 1]));
    tc_st16(c.tm + 64 + k0, hi);
    tc_st8(c.tm + 128 + k0 / 2, lo);
}
// layer 0's A row from the window's network state x9, dt and the four projected thrust inputs (K order tc_kmap0)
__device__ __forceinline__ void tc_put_inputs(const TcCtx& c, const float* x9, float dt, const float* u4) {
    const float z16[16] = {x9[0], x9[1], x9[2], x9[3], x9[4], x9[5], x9[6], x9[7], x9[8], dt, 0.0f, 0.0f,
                           u4[0], u4[1], u4[2], u4[3]};
    tc_put16(c, 0, z16);
}

// One layer's products, issued by ONE thread after the tile's threads have synchronised on the freshly written A.
// K, N are compile-time: the loop unrolls into 20 (hidden layers) tcgen05.mma with descriptors that differ by an
// add on the address field.
template <int K, int N>
__device__ __forceinline__ void tc_issue_layer(const TcCtx& c, int whi_off, int wlo_off, int w16_off) {
    constexpr uint32_t lbo = (uint32_t)(N / 8) * 128;      // K-chunk stride in bytes, both operand widths
    constexpr uint32_t id32 = tc_idesc(TC_M, N, true), id16 = tc_idesc(TC_M, N, false);
    constexpr uint64_t step = (uint64_t)((2 * lbo) >> 4);  // two K chunks per instruction, in descriptor address units
    const uint64_t dhi = tc_desc(tc_smem(c.sw + whi_off), lbo, 128), dlo = tc_desc(tc_smem(c.sw + wlo_off), lbo, 128),
                   d16 = tc_desc(tc_smem(c.sw + w16_off), lbo, 128);
#pragma unroll
    for (int k8 = 0; k8 < K / 8; ++k8) tc_mma_tf32(c.tm0, c.tm0 + 64 + 8 * k8, dhi + k8 * step, id32, k8 > 0);
#pragma unroll
    for (int k8 = 0; k8 < K / 8; ++k8) tc_mma_tf32(c.tm0, c.tm0 + 64 + 8 * k8, dlo + k8 * step, id32, true);
#pragma unroll
    for (int k16 = 0; k16 < K / 16; ++k16) tc_mma_f16(c.tm0, c.tm0 + 128 + 8 * k16, d16 + k16 * step, id16);
    tc_commit(c.bar);
}
// Starts one layer for the tile: every thread has stored its row of A; the tile's first thread issues the MMAs.
template <int K, int N>
__device__ __forceinline__ void tc_layer_start(const TcCtx& c, int whi_off, int wlo_off, int w16_off) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + c.tile), "r"(TC_M) : "memory");     // the tile's 128 threads
    if (c.row == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc_issue_layer<K, N>(c, whi_off, wlo_off, w16_off);
    }
}
// Returns with the layer's accumulators readable (and its A overwritable).
__device__ __forceinline__ void tc_layer_wait(TcCtx& c) {
    tc_bar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// bias + AdaptiveSoftplus + LayerNorm of one hidden layer (scales and affine folded, see the header): the row's 64
// accumulators -> the next layer's A
__device__ __forceinline__ void tc_hidden_epilogue(const TcCtx& c, int layer, const TcAct& act) {
    const float* par = c.sw + TC_PAR + layer * 64;
    const float sc = act.sc[layer];
    float a[HID];
    tc_ld32(c.tm, a);
    tc_ld32(c.tm + 32, a + 32);
    float m = 0.0f;
#pragma unroll
    for (int j = 0; j < HID; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(par + j);
        const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float t = fmaf(sc, a[j + q], bb[q]);
            float e, l;
            asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-fabsf(t)));
            asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(1.0f + e));
            a[j + q] = fmaxf(t, 0.0f) + l;
        }
        m += (a[j] + a[j + 1]) + (a[j + 2] + a[j + 3]);
    }
    m *= (1.0f / HID);
    float v0 = 0.0f, v1 = 0.0f;
#pragma unroll
    for (int j = 0; j < HID; j += 2) {
        a[j] -= m;
        a[j + 1] -= m;
        v0 = fmaf(a[j], a[j], v0);
        v1 = fmaf(a[j + 1], a[j + 1], v1);
    }
    const float rstd = act.sg[layer] * rsqrtf((v0 + v1) * (1.0f / HID) + act.eps[layer]);
    // the tile's MMAs of this layer are complete (every thread waited on the commit): A may be overwritten
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float out[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) out[j] = a[16 * q + j] * rstd;
        tc_put16(c, 16 * q, out);
    }
}

// The network of PINcNet.forward (training/train_tank_brov2_rk4.py:627-673) for the tile's 128 windows.  On entry
// every thread has stored its row of layer 0's A (tc_put_inputs); on return the OUTPUT layer's MMAs are in flight:
// call tc_layer_wait, then tc_residual.  Every thread of the tile must call it (tile-wide barriers inside).
__device__ __forceinline__ void pinc_net_tc(TcCtx& c, const TcAct& act) {
    tc_layer_start<16, HID>(c, TC_L0_HI, TC_L0_LO, TC_H16_L0);
    tc_layer_wait(c);
    tc_hidden_epilogue(c, 0, act);
#pragma unroll 1
    for (int l = 1; l <= 3; ++l) {
        tc_layer_start<HID, HID>(c, TC_L1 + (l - 1) * 8192, TC_L1 + (l - 1) * 8192 + 4096, TC_H16_L1 + (l - 1) * 2048);
        tc_layer_wait(c);
        tc_hidden_epilogue(c, l, act);
    }
    tc_layer_start<HID, 16>(c, TC_L4_HI, TC_L4_LO, TC_H16_L4);
}
// residual update from the output layer's accumulators: body-frame (dx, dy) rotated by the CURRENT yaw; (cos, sin)
// re-normalised (:639-673).  z = the 9 network states the step started from.
__device__ __forceinline__ void tc_residual(const TcCtx& c, const float* z, float (&xn)[9]) {
    float dx[16];
    tc_ld16(c.tm, dx);
    const float* b4 = c.sw + TC_PAR + 4 * 64;
#pragma unroll
    for (int j = 0; j < 9; ++j) dx[j] += b4[j];
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
}

// per-CTA setup / teardown (ALL threads of the CTA): stage the blob, init the tiles' mbarriers, allocate TMEM
// (warp 0 owns the allocation)
__device__ __forceinline__ void tc_setup(TcCtx& c, float* smem, const float* __restrict__ wtc, uint32_t* tmem_slot,
                                         uint64_t* bars) {
    c.sw = smem;
    for (int e = threadIdx.x * 4; e < TC_NW; e += blockDim.x * 4)
        *reinterpret_cast<float4*>(smem + e) = *reinterpret_cast<const float4*>(wtc + e);
    c.tile = threadIdx.x / TC_M;
    c.row = threadIdx.x % TC_M;
    c.bar = tc_smem(bars + c.tile);
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
    c.tm0 = *tmem_slot + (uint32_t)(c.tile * TC_COLS);
    c.tm = c.tm0 + ((uint32_t)(c.row & ~31) << 16);                  // a warp addresses the 32 lanes of its quarter
    c.phase = 0;
}
__device__ __forceinline__ void tc_teardown(const uint32_t* tmem_slot) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
}
