// Eval-mode PointNet trunk as ONE kernel per chain: cloud tile -> (xy transform) -> up to six point-wise
// layers (Conv1d k=1 with BatchNorm folded into weights and bias, ReLU) -> max over the cloud's points.
// No [rows, C] activation ever reaches HBM: a CTA keeps one superpoint (128 points = one UMMA M tile) on
// chip from the input tile to the pooled row.
//
//   reference: learning/pointnet.py:120-133 (PointNet.forward) and :55-61 (STNkD.forward) under
//   model.eval() (learning/main.py:229-311): conv -> BatchNorm1d(running statistics) -> ReLU chains, then
//   F.max_pool1d over the points.
//
// Per CTA (persistent over superpoints b = blockIdx.x, blockIdx.x + gridDim.x, ...):
//   warp 8 (one thread)  TMA producer: the superpoint's [F, 128] input tile (cp.async.bulk.tensor over the
//                        NCL clouds tensor) and the weight stream — for every layer, N-tile and 32-float
//                        K-chunk one [N_tile x 128 B] hi block and one lo block of the pre-split,
//                        pre-swizzled weight image (the weights live in L2; 2-stage ring of 32 KB)
//   warp 9 (one thread)  MMA issuer: tcgen05.mma kind::tf32, 3 MMAs per product (3xTF32: fp32-equivalent),
//                        A = the layer's input activations in shared memory (K-major SWIZZLE_128B hi/lo),
//                        B = the weight stage, D = TMEM accumulator (one 128-column tile per N-tile)
//   warps 0-7            one thread per (point, half of the channel blocks): input tile -> transform -> tf32 split -> A; after every
//                        layer tcgen05.ld of the accumulator row, + folded bias, ReLU, tf32 split, written
//                        IN PLACE as the next layer's A operand; after the last layer the max over the 128
//                        points (redux.sync on the non-negative float bits) -> pooled[b, :]
// mbarriers: w_full/w_empty (weight ring), x_full/x_empty (input double buffer), a_ready (A operand of the
// next layer is in shared memory), acc_full[t] (all MMAs of N-tile t have completed).
#include <cuda.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace spg {

constexpr int PF_ROWS = 128;          // points per superpoint = UMMA M
constexpr int PF_KC = 32;             // floats per K chunk (one 128-byte swizzle row)
constexpr int PF_MAXK = 128;          // widest layer input kept in shared memory
constexpr int PF_NT = 128;            // accumulator tile width (TMEM columns per N-tile).  (64-column tiles would let
                                      // the activation warps drain tile t under the MMAs of tile t+1, but the next A
                                      // operand is written IN PLACE: every MMA of the layer must have finished first)
constexpr int PF_MAX_TILES = 2;
constexpr int PF_STAGES_TMEM_A = 6;   // weight ring depth (32 KB stages); the A operand lives in tensor memory
constexpr int PF_STAGE_BYTES = 2 * PF_NT * PF_KC * 4;        // hi + lo of one (N-tile, K-chunk) = 32 KB
constexpr int PF_MAXF = 16;            // input features (S3DIS 14, Semantic3D 11, vKITTI 9)
constexpr int PF_X_BYTES = PF_MAXF * PF_ROWS * 4;            // one input tile buffer (8 KB)
constexpr int PF_MAX_LAYERS = 6;
constexpr int PF_ACT_WARPS = 8;        // activation warps: two per TMEM lane quarter (alternate 32-column blocks)
constexpr int PF_THREADS = (PF_ACT_WARPS + 2) * 32;
constexpr int PF_MAX_BIAS = 1024;

struct PfLayer {
    int K, N;      // K padded to a multiple of 32, N in {64, 128, 256}
    int w_row;     // first row of this layer's blocks in the weight image ([rows][32 floats])
    int b_off;     // offset of the folded bias in the bias vector
};

struct PfArgs {
    int n_layers;
    PfLayer L[PF_MAX_LAYERS];
    int F;                 // input features (<= 32)
    int64_t B;             // superpoints
    const float* T;        // [B, 4] spatial transformer output (xy' = xy (T + I)), or null
    int add_eye;
    const float* bias;     // folded biases of all layers
    int n_bias;
    float* pooled;         // [B, ldp]
    int64_t ldp;
};

__device__ __forceinline__ void pf_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void pf_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void pf_tma_2d(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(dst), "l"(map), "r"(x), "r"(y), "r"(bar)
        : "memory");
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
        "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
        "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}

// The A operand (the layer's input activations, tf32 hi | lo) lives in TENSOR MEMORY: the activation warps
// write the row they own with tcgen05.st (TMEM lane = point, column = channel — the layout the accumulator
// row already has, no shared-memory transposition), the MMA reads it directly, and all of shared memory goes
// to a 6-stage weight ring.  TMEM columns: [0,256) accumulators, [256,384) A hi, [384,512) A lo.
// (A shared-memory A operand — K-major SWIZZLE_128B, 128 KB, 2-stage ring — was measured 7 % slower.)
__global__ void __launch_bounds__(PF_THREADS, 1)
pointnet_fused_kernel(const PfArgs p, const __grid_constant__ CUtensorMap wmap,
                      const __grid_constant__ CUtensorMap xmap) {
    constexpr int PF_STAGES = PF_STAGES_TMEM_A;
    constexpr uint32_t TM_COLS = 512u;
    constexpr uint32_t TM_AHI = 256u, TM_ALO = 384u;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    // [weight ring 6 x 32 KB][input tiles 2 x 8 KB][bias][pool scratch]
    uint8_t* w_s = smem;
    uint8_t* x_s = w_s + PF_STAGES * PF_STAGE_BYTES;
    float* bias_s = reinterpret_cast<float*>(x_s + 2 * PF_X_BYTES);
    float* pool_s = bias_s + PF_MAX_BIAS;  // [4 warps][256]
    __shared__ __align__(8) uint64_t bars[2 * PF_STAGES_TMEM_A + 4 + 1 + PF_MAX_TILES];
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t bars_u = smem_u32(&bars[0]);
    auto w_full = [&](int s) { return bars_u + 8u * (uint32_t)s; };
    auto w_empty = [&](int s) { return bars_u + 8u * (uint32_t)(PF_STAGES + s); };
    auto x_full = [&](int s) { return bars_u + 8u * (uint32_t)(2 * PF_STAGES + s); };
    auto x_empty = [&](int s) { return bars_u + 8u * (uint32_t)(2 * PF_STAGES + 2 + s); };
    const uint32_t a_ready = bars_u + 8u * (2 * PF_STAGES + 4);
    auto acc_full = [&](int nt) { return bars_u + 8u * (uint32_t)(2 * PF_STAGES + 5 + nt); };

    if (t == 0) {
        for (int s = 0; s < PF_STAGES; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(x_full(s), 1);
            mbar_init(x_empty(s), PF_ACT_WARPS);
        }
        for (int s = 0; s < PF_MAX_TILES; ++s) mbar_init(acc_full(s), 1);
        mbar_init(a_ready, PF_ACT_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)), "r"(TM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    SPG_PDL_ENTRY();  // set-up above overlaps the previous kernel of the stream; global memory only below
    for (int i = t; i < p.n_bias; i += PF_THREADS) bias_s[i] = p.bias[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t w_u = smem_u32(w_s), x_u = smem_u32(x_s);

    if (warp == PF_ACT_WARPS) {
        // ================================ TMA producer ================================
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            uint32_t it = 0, ci = 0;
            for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++ci) {
                const int xb = ci & 1;
                mbar_wait(x_empty(xb), ((ci >> 1) & 1) ^ 1);
                pf_expect_tx(x_full(xb), (uint32_t)(p.F * PF_ROWS * 4));
                pf_tma_2d(x_u + xb * PF_X_BYTES, &xmap, 0, (int)(b * p.F), x_full(xb));
                for (int l = 0; l < p.n_layers; ++l) {
                    const PfLayer& Ly = p.L[l];
                    const int ntile = Ly.N < PF_NT ? Ly.N : PF_NT;
                    for (int nt = 0; nt < Ly.N / ntile; ++nt)
                        for (int kc = 0; kc < Ly.K / PF_KC; ++kc, ++it) {
                            const int s = it % PF_STAGES;
                            mbar_wait(w_empty(s), ((it / PF_STAGES) & 1) ^ 1);
                            pf_expect_tx(w_full(s), (uint32_t)(2 * ntile * PF_KC * 4));
                            const uint32_t dst = w_u + s * PF_STAGE_BYTES;
                            // image rows of this layer: [(kc*2 + half)*N + n]
                            for (int half = 0; half < 2; ++half)
                                for (int sub = 0; sub < ntile / 32; ++sub)  // TMA boxes of 32 rows
                                    pf_tma_2d(dst + (uint32_t)(half * ntile + sub * 32) * (PF_KC * 4), &wmap, 0,
                                              Ly.w_row + (kc * 2 + half) * Ly.N + nt * ntile + sub * 32, w_full(s));
                        }
                }
            }
        }
    } else if (warp == PF_ACT_WARPS + 1) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            uint32_t it = 0, q = 0;
            for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
                for (int l = 0; l < p.n_layers; ++l, ++q) {
                    const PfLayer& Ly = p.L[l];
                    const int ntile = Ly.N < PF_NT ? Ly.N : PF_NT;
                    const uint32_t idesc = umma_idesc_tf32(PF_ROWS, ntile);
                    mbar_wait(a_ready, q & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int nt = 0; nt < Ly.N / ntile; ++nt) {
                        const uint32_t d = tmem_base + (uint32_t)(nt * PF_NT);
                        for (int kc = 0; kc < Ly.K / PF_KC; ++kc, ++it) {
                            const int s = it % PF_STAGES;
                            mbar_wait(w_full(s), (it / PF_STAGES) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t b_hi = w_u + (uint32_t)s * PF_STAGE_BYTES;
                            const uint32_t b_lo = b_hi + (uint32_t)ntile * PF_KC * 4;
#pragma unroll
                            for (int ks = 0; ks < PF_KC / 8; ++ks) {
                                const uint32_t ko = ks * 32;
                                const uint64_t dbh = umma_desc_k_sw128(b_hi + ko), dbl = umma_desc_k_sw128(b_lo + ko);
                                const uint32_t th = tmem_base + TM_AHI + (uint32_t)(kc * PF_KC + ks * 8);
                                const uint32_t tl = tmem_base + TM_ALO + (uint32_t)(kc * PF_KC + ks * 8);
                                umma_tf32_ts(d, th, dbh, idesc, (kc | ks) ? 1u : 0u);
                                umma_tf32_ts(d, tl, dbh, idesc, 1u);
                                umma_tf32_ts(d, th, dbl, idesc, 1u);
                            }
                            umma_commit(w_empty(s));
                        }
                        umma_commit(acc_full(nt));
                    }
                }
            }
        }
    } else {
        // ======================= activation warps: one thread per point =======================
        const int quarter = warp & 3, half = warp >> 2;  // TMEM lane quarter; which 32-column blocks
        const int row = quarter * 32 + lane;             // 0..127 = TMEM lane = point of the superpoint
        uint32_t ci = 0, acc_cnt[PF_MAX_TILES] = {0u, 0u};
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++ci) {
            // ---- input tile -> (xy transform) -> A chunk 0
            const int xb = ci & 1;
            mbar_wait(x_full(xb), (ci >> 1) & 1);
            const float* xin = reinterpret_cast<const float*>(x_s + xb * PF_X_BYTES);
            float v[PF_KC];  // one K chunk: the features, zero-padded to 32
#pragma unroll
            for (int f = 0; f < PF_KC; ++f) v[f] = (f < PF_MAXF && f < p.F) ? xin[f * PF_ROWS + row] : 0.f;
            __syncwarp();
            if (lane == 0) pf_arrive(x_empty(xb));
            if (p.T && p.F >= 2) {
                const float eye = p.add_eye ? 1.f : 0.f;
                const float t00 = p.T[b * 4 + 0] + eye, t01 = p.T[b * 4 + 1], t10 = p.T[b * 4 + 2],
                            t11 = p.T[b * 4 + 3] + eye;
                const float x0 = v[0], x1 = v[1];
                v[0] = fmaf(x0, t00, x1 * t10);  // row vector times T (pointnet.py:123)
                v[1] = fmaf(x0, t01, x1 * t11);
            }
            const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
            auto store_chunk = [&](int kc, const float (&o)[32]) {
                // 32 consecutive channels of this thread's row = K chunk kc of the next A operand
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    hi[j] = to_tf32(o[j]);
                    lo[j] = to_tf32(o[j] - __uint_as_float(hi[j]));
                }
                tmem_st32(lane_base + TM_AHI + (uint32_t)(kc * PF_KC), hi);
                tmem_st32(lane_base + TM_ALO + (uint32_t)(kc * PF_KC), lo);
            };
            auto publish_a = [&]() {
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) pf_arrive(a_ready);
            };
            if (half == 0) store_chunk(0, v);
            publish_a();
            // ---- layers
            for (int l = 0; l < p.n_layers; ++l) {
                const PfLayer& Ly = p.L[l];
                const int ntile = Ly.N < PF_NT ? Ly.N : PF_NT;
                const bool last = l + 1 == p.n_layers;
                for (int nt = 0; nt < Ly.N / ntile; ++nt) {
                    mbar_wait(acc_full(nt), acc_cnt[nt] & 1);
                    ++acc_cnt[nt];
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int cb = half; cb < ntile / 32; cb += PF_ACT_WARPS / 4) {
                        uint32_t r[32];
                        tmem_ld32(lane_base + (uint32_t)(nt * PF_NT + cb * 32), r);
                        const int col0 = nt * ntile + cb * 32;
                        const float* bs = bias_s + Ly.b_off + col0;
                        if (!last) {
                            // next layer's A operand: these 32 columns are exactly K chunk col0/32
                            float o[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[j] = fmaxf(__uint_as_float(r[j]) + bs[j], 0.f);
                            store_chunk(col0 / PF_KC, o);
                        } else {
                            // max over the 32 points of this warp: ReLU output is >= 0, so the float order
                            // is the unsigned order of the bit patterns (one redux.sync per column)
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float o = fmaxf(__uint_as_float(r[j]) + bs[j], 0.f);
                                const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(o));
                                if (lane == j) pool_s[quarter * 256 + col0 + j] = __uint_as_float(m);
                            }
                        }
                    }
                }
                if (!last) {
                    // (all MMAs that read the old A have completed: acc_full of every N-tile was waited for)
                    publish_a();
                } else {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(PF_ACT_WARPS * 32) : "memory");
                    for (int c = t; c < Ly.N; c += PF_ACT_WARPS * 32) {
                        const float m = fmaxf(fmaxf(pool_s[c], pool_s[256 + c]), fmaxf(pool_s[512 + c], pool_s[768 + c]));
                        p.pooled[b * p.ldp + c] = m;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(PF_ACT_WARPS * 32) : "memory");
                }
            }
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TM_COLS) : "memory");
    }
}

// ------------------------------------------------------------------ bf16 variant (configs[3] arithmetic)
// Same trunk with bf16 operands and fp32 accumulation: tcgen05.mma kind::f16, ONE MMA per product.
// Activations travel between layers as bf16 in shared memory (K-major SWIZZLE_128B, 64 elements per
// 128-byte row, written by the thread that owns the row); the weight image is bf16 (a quarter of the
// 3xTF32 stream), 8-stage TMA ring.  Inputs and every layer output are rounded to bf16 (8-bit mantissa):
// results agree with the fp32 path to ~1e-2 (tests state the bound), not to 1e-4.
constexpr int PB_KC = 64;                                   // bf16 elements per K chunk (128-byte row)
constexpr int PB_STAGES = 8;
constexpr int PB_STAGE_BYTES = PF_NT * 128;                 // one (N-tile, K-chunk) weight block = 16 KB
constexpr int PB_A_CHUNK_BYTES = PF_ROWS * 128;             // 16 KB
constexpr int PB_A_BYTES = (PF_MAXK / PB_KC) * PB_A_CHUNK_BYTES;  // 32 KB

__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N) {
    // c = F32 (1 << 4), a = b = BF16 (1 << 7, 1 << 10), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ uint32_t pack_bf16x2(float e0, float e1) {  // e0 -> low half (lower address)
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(e1), "f"(e0));
    return r;
}

__global__ void __launch_bounds__(PF_THREADS, 1)
pointnet_fused_bf16_kernel(const PfArgs p, const __grid_constant__ CUtensorMap wmap,
                           const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    // [A operand 32 KB][weight ring 8 x 16 KB][input tiles 2 x 8 KB][bias][pool scratch]
    uint8_t* a_s = smem;
    uint8_t* w_s = a_s + PB_A_BYTES;
    uint8_t* x_s = w_s + PB_STAGES * PB_STAGE_BYTES;
    float* bias_s = reinterpret_cast<float*>(x_s + 2 * PF_X_BYTES);
    float* pool_s = bias_s + PF_MAX_BIAS;
    __shared__ __align__(8) uint64_t bars[2 * PB_STAGES + 4 + 1 + PF_MAX_TILES];
    __shared__ uint32_t tmem_base_s;
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const uint32_t bars_u = smem_u32(&bars[0]);
    auto w_full = [&](int s) { return bars_u + 8u * (uint32_t)s; };
    auto w_empty = [&](int s) { return bars_u + 8u * (uint32_t)(PB_STAGES + s); };
    auto x_full = [&](int s) { return bars_u + 8u * (uint32_t)(2 * PB_STAGES + s); };
    auto x_empty = [&](int s) { return bars_u + 8u * (uint32_t)(2 * PB_STAGES + 2 + s); };
    const uint32_t a_ready = bars_u + 8u * (2 * PB_STAGES + 4);
    auto acc_full = [&](int nt) { return bars_u + 8u * (uint32_t)(2 * PB_STAGES + 5 + nt); };
    if (t == 0) {
        for (int s = 0; s < PB_STAGES; ++s) {
            mbar_init(w_full(s), 1);
            mbar_init(w_empty(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(x_full(s), 1);
            mbar_init(x_empty(s), PF_ACT_WARPS);
        }
        for (int s = 0; s < PF_MAX_TILES; ++s) mbar_init(acc_full(s), 1);
        mbar_init(a_ready, PF_ACT_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    SPG_PDL_ENTRY();  // set-up above overlaps the previous kernel of the stream; global memory only below
    for (int i = t; i < p.n_bias; i += PF_THREADS) bias_s[i] = p.bias[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_u = smem_u32(a_s), w_u = smem_u32(w_s), x_u = smem_u32(x_s);

    if (warp == PF_ACT_WARPS) {
        if (lane == 0) {  // ---- TMA producer
            asm volatile("prefetch.tensormap [%0];" ::"l"(&wmap) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&xmap) : "memory");
            uint32_t it = 0, ci = 0;
            for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++ci) {
                const int xb = ci & 1;
                mbar_wait(x_empty(xb), ((ci >> 1) & 1) ^ 1);
                pf_expect_tx(x_full(xb), (uint32_t)(p.F * PF_ROWS * 4));
                pf_tma_2d(x_u + xb * PF_X_BYTES, &xmap, 0, (int)(b * p.F), x_full(xb));
                for (int l = 0; l < p.n_layers; ++l) {
                    const PfLayer& Ly = p.L[l];
                    const int ntile = Ly.N < PF_NT ? Ly.N : PF_NT;
                    for (int nt = 0; nt < Ly.N / ntile; ++nt)
                        for (int kc = 0; kc < Ly.K / PB_KC; ++kc, ++it) {
                            const int s = it % PB_STAGES;
                            mbar_wait(w_empty(s), ((it / PB_STAGES) & 1) ^ 1);
                            pf_expect_tx(w_full(s), (uint32_t)(ntile * 128));
                            // image rows of this layer: [kc*N + n], 128 bytes each; boxes of 32 rows
                            for (int sub = 0; sub < ntile / 32; ++sub)
                                pf_tma_2d(w_u + (uint32_t)(s * PB_STAGE_BYTES + sub * 32 * 128), &wmap, 0,
                                          Ly.w_row + kc * Ly.N + nt * ntile + sub * 32, w_full(s));
                        }
                }
            }
        }
    } else if (warp == PF_ACT_WARPS + 1) {
        if (lane == 0) {  // ---- MMA issuer
            uint32_t it = 0, q = 0;
            for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x) {
                for (int l = 0; l < p.n_layers; ++l, ++q) {
                    const PfLayer& Ly = p.L[l];
                    const int ntile = Ly.N < PF_NT ? Ly.N : PF_NT;
                    const uint32_t idesc = umma_idesc_bf16(PF_ROWS, ntile);
                    mbar_wait(a_ready, q & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int nt = 0; nt < Ly.N / ntile; ++nt) {
                        const uint32_t d = tmem_base + (uint32_t)(nt * PF_NT);
                        for (int kc = 0; kc < Ly.K / PB_KC; ++kc, ++it) {
                            const int s = it % PB_STAGES;
                            mbar_wait(w_full(s), (it / PB_STAGES) & 1);
                            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                            const uint32_t a0 = a_u + (uint32_t)kc * PB_A_CHUNK_BYTES;
                            const uint32_t b0 = w_u + (uint32_t)s * PB_STAGE_BYTES;
#pragma unroll
                            for (int ks = 0; ks < PB_KC / 16; ++ks)  // K = 16 bf16 = 32 bytes per instruction
                                umma_bf16(d, umma_desc_k_sw128(a0 + ks * 32), umma_desc_k_sw128(b0 + ks * 32), idesc,
                                          (kc | ks) ? 1u : 0u);
                            umma_commit(w_empty(s));
                        }
                        umma_commit(acc_full(nt));
                    }
                }
            }
        }
    } else {
        // ---- activation warps: one thread per (point, half of the 32-column blocks)
        const int quarter = warp & 3, half = warp >> 2;
        const int row = quarter * 32 + lane;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quarter * 32) << 16);
        // 32 consecutive channels starting at channel col0 of the next A operand: 4 x 16 bytes of packed bf16
        auto store_block = [&](int col0, const float (&o)[32]) {
            const uint32_t base = a_u + (uint32_t)(col0 / PB_KC) * PB_A_CHUNK_BYTES;
            const int c16_0 = (col0 % PB_KC) / 8;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t q0 = pack_bf16x2(o[8 * j], o[8 * j + 1]), q1 = pack_bf16x2(o[8 * j + 2], o[8 * j + 3]);
                const uint32_t q2 = pack_bf16x2(o[8 * j + 4], o[8 * j + 5]), q3 = pack_bf16x2(o[8 * j + 6], o[8 * j + 7]);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(base + sw128_off(row, c16_0 + j)),
                             "r"(q0), "r"(q1), "r"(q2), "r"(q3) : "memory");
            }
        };
        auto publish_a = [&]() {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) pf_arrive(a_ready);
        };
        const float zeros[32] = {0.f};
        uint32_t ci = 0, acc_cnt[PF_MAX_TILES] = {0u, 0u};
        for (int64_t b = blockIdx.x; b < p.B; b += gridDim.x, ++ci) {
            const int xb = ci & 1;
            mbar_wait(x_full(xb), (ci >> 1) & 1);
            const float* xin = reinterpret_cast<const float*>(x_s + xb * PF_X_BYTES);
            float v[32];
#pragma unroll
            for (int f = 0; f < 32; ++f) v[f] = (f < PF_MAXF && f < p.F) ? xin[f * PF_ROWS + row] : 0.f;
            __syncwarp();
            if (lane == 0) pf_arrive(x_empty(xb));
            if (p.T && p.F >= 2) {
                const float eye = p.add_eye ? 1.f : 0.f;
                const float t00 = p.T[b * 4 + 0] + eye, t01 = p.T[b * 4 + 1], t10 = p.T[b * 4 + 2],
                            t11 = p.T[b * 4 + 3] + eye;
                const float x0 = v[0], x1 = v[1];
                v[0] = fmaf(x0, t00, x1 * t10);
                v[1] = fmaf(x0, t01, x1 * t11);
            }
            if (half == 0) store_block(0, v); else store_block(32, zeros);  // chunk 0: features | zero padding
            publish_a();
            for (int l = 0; l < p.n_layers; ++l) {
                const PfLayer& Ly = p.L[l];
                const int ntile = Ly.N < PF_NT ? Ly.N : PF_NT;
                const bool last = l + 1 == p.n_layers;
                for (int nt = 0; nt < Ly.N / ntile; ++nt) {
                    mbar_wait(acc_full(nt), acc_cnt[nt] & 1);
                    ++acc_cnt[nt];
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int cb = half; cb < ntile / 32; cb += PF_ACT_WARPS / 4) {
                        uint32_t r[32];
                        tmem_ld32(lane_base + (uint32_t)(nt * PF_NT + cb * 32), r);
                        const int col0 = nt * ntile + cb * 32;
                        const float* bs = bias_s + Ly.b_off + col0;
                        if (!last) {
                            float o[32];
#pragma unroll
                            for (int j = 0; j < 32; ++j) o[j] = fmaxf(__uint_as_float(r[j]) + bs[j], 0.f);
                            store_block(col0, o);
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float o = fmaxf(__uint_as_float(r[j]) + bs[j], 0.f);
                                const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(o));
                                if (lane == j) pool_s[quarter * 256 + col0 + j] = __uint_as_float(m);
                            }
                        }
                    }
                    // a 32-wide layer fills only half of the next 64-element K chunk: zero the other half
                    if (!last && Ly.N == 32 && half == 1) store_block(32, zeros);
                }
                if (!last) {
                    publish_a();
                } else {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("bar.sync 1, %0;" ::"n"(PF_ACT_WARPS * 32) : "memory");
                    for (int c = t; c < Ly.N; c += PF_ACT_WARPS * 32) {
                        const float m = fmaxf(fmaxf(pool_s[c], pool_s[256 + c]), fmaxf(pool_s[512 + c], pool_s[768 + c]));
                        p.pooled[b * p.ldp + c] = m;
                    }
                    asm volatile("bar.sync 1, %0;" ::"n"(PF_ACT_WARPS * 32) : "memory");
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// bf16 weight image of diag(row_scale) * W: for every 64-element K chunk [N rows][128 B], SWIZZLE_128B
__global__ void pack_weights_bf16_kernel(const float* __restrict__ W, int64_t ldw,
                                         const float* __restrict__ row_scale, int N, int K, int k_valid,
                                         uint16_t* __restrict__ img) {
    SPG_PDL_ENTRY();
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    float v = 0.f;
    if (k < k_valid) v = W[(int64_t)n * ldw + k] * (row_scale ? row_scale[n] : 1.f);
    const uint32_t pk = pack_bf16x2(v, 0.f);
    const int kc = k / PB_KC, kk = k % PB_KC;
    img[(int64_t)kc * N * PB_KC + (sw128_off(n, kk >> 3) >> 1) + (kk & 7)] = (uint16_t)(pk & 0xffffu);
}

typedef CUresult (*PfEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                               const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                               CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PfEncodeFn pf_encode() {
    static PfEncodeFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PfEncodeFn>(ptr);
    });
    return fn;
}

static int pf_map_2d(CUtensorMap* map, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                     uint32_t box_rows, bool bf16 = false) {
    PfEncodeFn fn = pf_encode();
    if (!fn) return SPG_E_UNSUPPORTED;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * (bf16 ? 2 : 4)};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                          const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? SPG_OK : SPG_E_BADARG;
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_pointnet_fused_supported(int n_features, int n_points, int n_layers, const int32_t* widths) {
    if (n_points != PF_ROWS || n_features < 1 || n_features > PF_MAXF) return 0;
    if (n_layers < 1 || n_layers > PF_MAX_LAYERS || !widths) return 0;
    int total = 0;
    for (int l = 0; l < n_layers; ++l) {
        const int n = widths[l];
        if (n != 32 && n != 64 && n != 128 && n != 256) return 0;
        if (l + 1 < n_layers && n > PF_MAXK) return 0;  // a layer's output is the next layer's K
        total += n;
    }
    return total <= PF_MAX_BIAS ? 1 : 0;
}

int64_t spg_pointnet_fused_image_rows(int n_features, int n_layers, const int32_t* widths) {
    (void)n_features;
    int64_t rows = 0;
    int k = PF_KC;  // first layer: features padded to one chunk
    for (int l = 0; l < n_layers; ++l) {
        rows += (int64_t)(k / PF_KC) * 2 * widths[l];
        k = widths[l];
    }
    return rows;
}

int spg_pointnet_fused_eval(const float* clouds, int64_t n_clouds, int n_features, int n_points, const float* T,
                            int add_eye, const float* weight_image, const float* bias, int n_layers,
                            const int32_t* widths, float* pooled, int64_t ldp, spg_stream_t stream) {
    if (n_clouds < 0 || !widths) return SPG_E_BADARG;
    if (!spg_pointnet_fused_supported(n_features, n_points, n_layers, widths)) return SPG_E_UNSUPPORTED;
    if (n_clouds == 0) return SPG_OK;
    if (!clouds || !weight_image || !bias || !pooled || ldp < widths[n_layers - 1]) return SPG_E_BADARG;
    if (((uintptr_t)clouds | (uintptr_t)weight_image) & 15) return SPG_E_ALIGN;
    if (n_clouds * n_features >= (1ll << 31)) return SPG_E_UNSUPPORTED;
    PfArgs a;
    a.n_layers = n_layers;
    int k = PF_KC, row = 0, boff = 0;
    for (int l = 0; l < n_layers; ++l) {
        a.L[l].K = k;
        a.L[l].N = widths[l];
        a.L[l].w_row = row;
        a.L[l].b_off = boff;
        row += (k / PF_KC) * 2 * widths[l];
        boff += widths[l];
        k = widths[l];
    }
    a.F = n_features; a.B = n_clouds; a.T = T; a.add_eye = add_eye; a.bias = bias; a.n_bias = boff;
    a.pooled = pooled; a.ldp = ldp;
    CUtensorMap wmap, xmap;
    int rc = pf_map_2d(&wmap, weight_image, PF_KC, (uint64_t)row, PF_KC, 32);
    if (rc) return rc;
    rc = pf_map_2d(&xmap, clouds, PF_ROWS, (uint64_t)(n_clouds * n_features), PF_ROWS, (uint32_t)n_features);
    if (rc) return rc;
    const int64_t grid = n_clouds < kNumSMs ? n_clouds : kNumSMs;
    const int smem = PF_STAGES_TMEM_A * PF_STAGE_BYTES + 2 * PF_X_BYTES + (PF_MAX_BIAS + 4 * 256) * 4;
    cudaError_t e = cudaFuncSetAttribute(pointnet_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    SPG_LAUNCH(K_POINTNET_FUSED, (cudaStream_t)stream, pointnet_fused_kernel, (unsigned)grid, PF_THREADS, smem,
               a, wmap, xmap);
    return launch_status();
}

/* bf16 twins (spg_b200.h): K chunks of 64 elements; a 32-wide layer is followed by a zero-padded chunk */
int64_t spg_pointnet_fused_bf16_image_rows(int n_features, int n_layers, const int32_t* widths) {
    (void)n_features;
    int64_t rows = 0;
    int k = PB_KC;
    for (int l = 0; l < n_layers; ++l) {
        rows += (int64_t)(k / PB_KC) * widths[l];
        k = widths[l] < PB_KC ? PB_KC : widths[l];
    }
    return rows;
}

int spg_tc_pack_weights_bf16(const float* W, int64_t ldw, const float* row_scale, int N, int K, int k_valid,
                             void* image, spg_stream_t stream) {
    if (!W || !image || N <= 0 || K <= 0 || k_valid <= 0 || k_valid > K) return SPG_E_BADARG;
    if (K % PB_KC != 0 || N % 8 != 0) return SPG_E_UNSUPPORTED;
    const int64_t total = (int64_t)N * K;
    SPG_LAUNCH(K_TC_PACK, (cudaStream_t)stream, pack_weights_bf16_kernel, (unsigned)ceil_div64(total, 256), 256, 0, W,
               ldw, row_scale, N, K, k_valid, (uint16_t*)image);
    return launch_status();
}

int spg_pointnet_fused_eval_bf16(const float* clouds, int64_t n_clouds, int n_features, int n_points, const float* T,
                                 int add_eye, const void* weight_image, const float* bias, int n_layers,
                                 const int32_t* widths, float* pooled, int64_t ldp, spg_stream_t stream) {
    if (n_clouds < 0 || !widths) return SPG_E_BADARG;
    if (!spg_pointnet_fused_supported(n_features, n_points, n_layers, widths)) return SPG_E_UNSUPPORTED;
    if (n_clouds == 0) return SPG_OK;
    if (!clouds || !weight_image || !bias || !pooled || ldp < widths[n_layers - 1]) return SPG_E_BADARG;
    if (((uintptr_t)clouds | (uintptr_t)weight_image) & 15) return SPG_E_ALIGN;
    if (n_clouds * n_features >= (1ll << 31)) return SPG_E_UNSUPPORTED;
    PfArgs a;
    a.n_layers = n_layers;
    int k = PB_KC, row = 0, boff = 0;
    for (int l = 0; l < n_layers; ++l) {
        a.L[l].K = k;
        a.L[l].N = widths[l];
        a.L[l].w_row = row;
        a.L[l].b_off = boff;
        row += (k / PB_KC) * widths[l];
        boff += widths[l];
        k = widths[l] < PB_KC ? PB_KC : widths[l];
    }
    a.F = n_features; a.B = n_clouds; a.T = T; a.add_eye = add_eye; a.bias = bias; a.n_bias = boff;
    a.pooled = pooled; a.ldp = ldp;
    CUtensorMap wmap, xmap;
    int rc = pf_map_2d(&wmap, weight_image, PB_KC, (uint64_t)row, PB_KC, 32, true);
    if (rc) return rc;
    rc = pf_map_2d(&xmap, clouds, PF_ROWS, (uint64_t)(n_clouds * n_features), PF_ROWS, (uint32_t)n_features);
    if (rc) return rc;
    const int64_t grid = n_clouds < kNumSMs ? n_clouds : kNumSMs;
    const int smem = PB_A_BYTES + PB_STAGES * PB_STAGE_BYTES + 2 * PF_X_BYTES + (PF_MAX_BIAS + 4 * 256) * 4;
    cudaError_t e = cudaFuncSetAttribute(pointnet_fused_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    SPG_LAUNCH(K_POINTNET_FUSED, (cudaStream_t)stream, pointnet_fused_bf16_kernel, (unsigned)grid, PF_THREADS, smem, a,
               wmap, xmap);
    return launch_status();
}

}  // extern "C"
