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
// A CTA is TWO independent tiles of 256 threads, each with its own mbarrier, named barriers and TMEM columns: while
// one tile's MMAs run (or wait to be issued), the other tile's epilogue keeps the CUDA cores and MUFU busy.  A row
// (window) of a tile is shared by TWO threads in warps w and w + 4 (the same TMEM lane quarter): each does the
// epilogue of 32 of the 64 columns, and the two exchange their partial LayerNorm statistics (mean and centred sum of
// squares, combined exactly) through shared memory — 16 warps per SM hide the MUFU / tcgen05.ld latencies that 4 or 8
// could not (measured on the 1M-window table: 1 tile x 128 threads, A in shared memory 73 ms; 2 tiles x 128 threads,
// A in TMEM 45 ms; the CUDA-core kernel 79 ms).  The first thread of a row ("leader") owns the window's network
// state and the scoring; the second ("helper") owns the thruster map (fp64 lag recursion), which it evaluates for the
// NEXT step while the output layer's MMAs of the current one run.
// Shared memory per CTA: weights 115 KB (hi + lo of five layers) + parameters 3 KB, one CTA per SM, persistent over
// tiles.  TMEM per tile: 64 columns of accumulators + 64 + 64 columns of A_hi / A_lo = 192; 512 allocated per CTA.
// What the epilogue does NOT compute (folded on the host, brov_pinc_create): the LayerNorm affine (weight, bias) of
// layer l is multiplied into layer l + 1's weights and bias in double precision before the hi/lo split; the
// activation's 1/beta and ln 2 factors are dropped because LayerNorm is scale-invariant up to its epsilon, which is
// rescaled instead; beta log2(e) multiplies the bias once.  AdaptiveSoftplus then is max(t, 0) + log2(1 + 2^-|t|) with
// t = fma(beta log2 e, acc, bias'): two MUFU and ~13 other instructions per activation including the split
// (a cvt.rna.tf32.f32 per part is emulated with 4 instructions on sm_100a; here hi = (bits + 0x1000) & ~0x1fff,
// lo = x - hi exactly, rounded the same way by its own + 0x1000 and the tensor core's truncation of the low 13 bits).
// Layer 0's K axis is permuted so that each owner writes whole 4-column groups: k 0..8 = network state, 9 = dt,
// 10..11 = 0, 12..15 = the four projected thrust inputs (the host permutes W0 alike).
#pragma once

constexpr int TC_M = 128;                       // windows per tile = threads per CTA = TMEM lanes
constexpr int TC_L0_HI = 0, TC_L0_LO = 1024;    // [64 n][16 k] each, core-matrix layout (K = 14 padded to 16)
constexpr int TC_L1 = 2048;                     // hidden layers l = 1..3: hi at TC_L1 + (l-1)*8192, lo 4096 later
constexpr int TC_L4_HI = TC_L1 + 3 * 8192, TC_L4_LO = TC_L4_HI + 1024;   // [16 n][64 k] each (9 outputs padded to 16)
constexpr int TC_PAR = TC_L4_LO + 1024;         // per hidden layer: folded bias[64] (log2 domain); then b4'[16]
constexpr int TC_NW = TC_PAR + 4 * 64 + 16;     // floats in the blob
// per-layer scalars of the folded activation + LayerNorm (host: brov_pinc_create)
struct TcAct {
    float sc[4];    // beta * log2(e)
    float eps[4];   // 1e-5 * ((beta + 1e-12) / ln 2)^2: LayerNorm's epsilon in the un-scaled activation's units
    float sg[4];    // sign of ln 2 / (beta + 1e-12)
};
constexpr int TC_TILES = 2;                     // tiles per CTA
constexpr int TC_TPT = 2 * TC_M;                // threads per tile: two per row
constexpr int TC_THREADS = TC_TILES * TC_TPT;   // 512
constexpr int TC_COLS = 192;                    // TMEM columns per tile: D [0,64), A_hi [64,128), A_lo [128,192)
constexpr size_t TC_SMEM_BYTES = (size_t)TC_NW * sizeof(float) + 64;
// input index (PINcNet: 9 state, 4 thrust, dt) held by column k of layer 0's A; -1 = zero padding
__host__ __device__ constexpr int tc_kmap0(int k) { return k < 9 ? k : (k == 9 ? 13 : (k < 12 ? -1 : k - 3)); }

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
// x = hi + lo, hi a TF32 number (nearest, ties away; finite inputs), lo the exact remainder pre-biased by half a TF32
// ulp so that the tensor core's truncation of its low 13 bits rounds it to nearest
__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
    hi = __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
    lo = __uint_as_float(__float_as_uint(x - hi) + 0x1000u);
}
// consecutive columns of this thread's TMEM lane <- registers
__device__ __forceinline__ void tc_st4(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v[0])),
                 "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3]))
                 : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
                 : "memory");
}
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

struct TcCtx {
    const float* sw;   // weights + parameters in shared memory (TC_NW floats)
    float2* xch;       // this tile's LayerNorm exchange slots [2 parts][TC_M rows]
    uint32_t bar;      // this tile's mbarrier (shared-window address): the MMA commits arrive on it
    uint32_t tm;       // this thread's TMEM address of the tile's column 0 (lane bits = first lane of its warp)
    uint32_t tm0;      // the tile's column 0 at lane 0 (what the MMA instructions address)
    uint32_t phase;    // parity the next wait expects
    int row;           // this thread's row of the tile = its window
    int tile;          // tile of the CTA this thread belongs to
    int part;          // 0: leader (columns 0..31 of the hidden layers), 1: helper (columns 32..63)
};

// N (4, 8 or 16) activations -> their TF32 hi / lo parts -> columns [k0, k0 + N) of this thread's row of A_hi / A_lo
template <int N>
__device__ __forceinline__ void tc_put(const TcCtx& c, int k0, const float* x) {
    float hi[N], lo[N];
#pragma unroll
    for (int j = 0; j < N; ++j) tc_split(x[j], hi[j], lo[j]);
    if constexpr (N == 4) {
        tc_st4(c.tm + 64 + k0, hi);
        tc_st4(c.tm + 128 + k0, lo);
    } else if constexpr (N == 8) {
        tc_st8(c.tm + 64 + k0, hi);
        tc_st8(c.tm + 128 + k0, lo);
    } else {
        tc_st16(c.tm + 64 + k0, hi);
        tc_st16(c.tm + 128 + k0, lo);
    }
}
// layer 0's A row: the leader's part (network state x9 and dt) and the thrust inputs' part (tc_kmap0)
__device__ __forceinline__ void tc_put_state(const TcCtx& c, const float* x9, float dt) {
    tc_put<8>(c, 0, x9);
    const float t[4] = {x9[8], dt, 0.0f, 0.0f};
    tc_put<4>(c, 8, t);
}
__device__ __forceinline__ void tc_put_thrust(const TcCtx& c, const float* u4) { tc_put<4>(c, 12, u4); }

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

// Starts one layer for the tile: every thread has stored its part of A; the tile's first thread issues the MMAs.
__device__ __forceinline__ void tc_layer_start(const TcCtx& c, int whi_off, int wlo_off, int K, int N) {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    asm volatile("bar.sync %0, %1;" ::"r"(1 + c.tile), "r"(TC_TPT) : "memory");     // the tile's 256 threads
    if (c.row == 0 && c.part == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tc_issue_layer(c, whi_off, wlo_off, K, N);
    }
}
// Returns with the layer's accumulators readable (and its A overwritable).
__device__ __forceinline__ void tc_layer_wait(TcCtx& c) {
    tc_bar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// bias + AdaptiveSoftplus + LayerNorm of one hidden layer (scales and affine folded, see the header): this thread's 32
// of the row's 64 accumulators; the row's two threads combine their (mean, centred sum of squares) exactly (Chan et
// al.); the result is split and stored as this thread's 32 columns of the next layer's A.
__device__ __forceinline__ void tc_hidden_epilogue(const TcCtx& c, int layer, const TcAct& act) {
    const float* par = c.sw + TC_PAR + layer * 64 + 32 * c.part;
    const float sc = act.sc[layer];
    float a[32];
    tc_ld32(c.tm + 32 * c.part, a);
    float m = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
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
    m *= (1.0f / 32);
    float m2 = 0.0f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        a[j] -= m;
        m2 = fmaf(a[j], a[j], m2);
    }
    c.xch[c.part * TC_M + c.row] = make_float2(m, m2);
    asm volatile("bar.sync %0, %1;" ::"r"(1 + TC_TILES + c.tile), "r"(TC_TPT) : "memory");
    const float2 o = c.xch[(c.part ^ 1) * TC_M + c.row];
    // (the slot is rewritten one layer later, after the tile-wide barrier of tc_layer_start: no second barrier here)
    const float d = 0.5f * (m - o.x);                                       // this half's mean - the row's mean
    const float rstd = act.sg[layer] * rsqrtf(fmaf(d, d, (m2 + o.y) * (1.0f / HID)) + act.eps[layer]);
    const float dr = d * rstd;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        float out[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) out[j] = fmaf(a[16 * q + j], rstd, dr);
        tc_put<16>(c, 32 * c.part + 16 * q, out);
    }
}

// The network of PINcNet.forward (training/train_tank_brov2_rk4.py:627-673) for the tile's 128 windows.  On entry
// both threads of a row have stored layer 0's A (tc_put_state / tc_put_thrust); on return the OUTPUT layer's MMAs are
// in flight: call tc_layer_wait, then the leader reads the increments with tc_residual.  Every thread of the tile
// must call it (tile-wide barriers inside).
__device__ __forceinline__ void pinc_net_tc(TcCtx& c, const TcAct& act) {
    tc_layer_start(c, TC_L0_HI, TC_L0_LO, 16, HID);
    tc_layer_wait(c);
    tc_hidden_epilogue(c, 0, act);
#pragma unroll 1
    for (int l = 1; l <= 3; ++l) {
        tc_layer_start(c, TC_L1 + (l - 1) * 8192, TC_L1 + (l - 1) * 8192 + 4096, HID, HID);
        tc_layer_wait(c);
        tc_hidden_epilogue(c, l, act);
    }
    tc_layer_start(c, TC_L4_HI, TC_L4_LO, HID, 16);
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

// per-CTA setup / teardown (ALL threads of the CTA): stage the weights, init the tiles' mbarriers, allocate TMEM
// (warp 0 owns the allocation)
__device__ __forceinline__ void tc_setup(TcCtx& c, float* smem, const float* __restrict__ wtc, uint32_t* tmem_slot,
                                         uint64_t* bars, float2* xch) {
    c.sw = smem;
    for (int e = threadIdx.x * 4; e < TC_NW; e += blockDim.x * 4)
        *reinterpret_cast<float4*>(smem + e) = *reinterpret_cast<const float4*>(wtc + e);
    c.tile = threadIdx.x / TC_TPT;
    c.part = (threadIdx.x % TC_TPT) / TC_M;
    c.row = threadIdx.x % TC_M;
    c.xch = xch + c.tile * 2 * TC_M;
    c.bar = tc_smem(bars + c.tile);
    if (c.row == 0 && c.part == 0) tc_bar_init(c.bar);
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
__device__ __forceinline__ void tc_teardown(const TcCtx& c, const uint32_t* tmem_slot) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(512u) : "memory");
}
