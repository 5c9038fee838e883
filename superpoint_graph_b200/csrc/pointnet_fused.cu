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
//   warp 4 (one thread)  TMA producer: the superpoint's [F, 128] input tile (cp.async.bulk.tensor over the
//                        NCL clouds tensor) and the weight stream — for every layer, N-tile and 32-float
//                        K-chunk one [N_tile x 128 B] hi block and one lo block of the pre-split,
//                        pre-swizzled weight image (the weights live in L2; 2-stage ring of 32 KB)
//   warp 5 (one thread)  MMA issuer: tcgen05.mma kind::tf32, 3 MMAs per product (3xTF32: fp32-equivalent),
//                        A = the layer's input activations in shared memory (K-major SWIZZLE_128B hi/lo),
//                        B = the weight stage, D = TMEM accumulator (one 128-column tile per N-tile)
//   warps 0-3            one thread per point: input tile -> transform -> tf32 split -> A; after every
//                        layer tcgen05.ld of the accumulator row, + folded bias, ReLU, tf32 split, written
//                        IN PLACE as the next layer's A operand; after the last layer the max over the 128
//                        points (redux.sync on the non-negative float bits) -> pooled[b, :]
// mbarriers: w_full/w_empty (weight ring), x_full/x_empty (input double buffer), a_ready (A operand of the
// next layer is in shared memory), acc_full[t] (all MMAs of N-tile t have completed).
#include <cuda.h>

#include <mutex>

#include "common.cuh"
#include "tc_common.cuh"

namespace spg {

constexpr int PF_ROWS = 128;          // points per superpoint = UMMA M
constexpr int PF_KC = 32;             // floats per K chunk (one 128-byte swizzle row)
constexpr int PF_MAXK = 128;          // widest layer input kept in shared memory
constexpr int PF_NT = 128;            // accumulator tile width (TMEM columns per N-tile)
constexpr int PF_STAGES = 2;
constexpr int PF_STAGE_BYTES = 2 * PF_NT * PF_KC * 4;        // hi + lo = 32 KB
constexpr int PF_A_CHUNK_BYTES = 2 * PF_ROWS * PF_KC * 4;    // hi + lo of one K chunk = 32 KB
constexpr int PF_A_BYTES = (PF_MAXK / PF_KC) * PF_A_CHUNK_BYTES;  // 128 KB
constexpr int PF_MAXF = 16;            // input features (S3DIS 14, Semantic3D 11, vKITTI 9)
constexpr int PF_X_BYTES = PF_MAXF * PF_ROWS * 4;            // one input tile buffer (8 KB)
constexpr int PF_MAX_LAYERS = 6;
constexpr int PF_THREADS = 192;
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

__device__ __forceinline__ void pf_split_store(uint32_t a_hi, int row, int c16, float4 v) {
    uint4 hi, lo;
    hi.x = to_tf32(v.x); hi.y = to_tf32(v.y); hi.z = to_tf32(v.z); hi.w = to_tf32(v.w);
    lo.x = to_tf32(v.x - __uint_as_float(hi.x));
    lo.y = to_tf32(v.y - __uint_as_float(hi.y));
    lo.z = to_tf32(v.z - __uint_as_float(hi.z));
    lo.w = to_tf32(v.w - __uint_as_float(hi.w));
    const uint32_t off = sw128_off(row, c16);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + off), "r"(hi.x), "r"(hi.y), "r"(hi.z),
                 "r"(hi.w) : "memory");
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a_hi + PF_ROWS * PF_KC * 4 + off), "r"(lo.x),
                 "r"(lo.y), "r"(lo.z), "r"(lo.w) : "memory");
}

__global__ void __launch_bounds__(PF_THREADS, 1)
pointnet_fused_kernel(const PfArgs p, const __grid_constant__ CUtensorMap wmap,
                      const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0u) __trap();
    // [A operand 128 KB][weight ring 2 x 32 KB][input tiles 2 x 8 KB][bias][pool scratch]
    uint8_t* a_s = smem;
    uint8_t* w_s = a_s + PF_A_BYTES;
    uint8_t* x_s = w_s + PF_STAGES * PF_STAGE_BYTES;
    float* bias_s = reinterpret_cast<float*>(x_s + 2 * PF_X_BYTES);
    float* pool_s = bias_s + PF_MAX_BIAS;  // [4 warps][256]
    __shared__ __align__(8) uint64_t bars[2 * PF_STAGES + 4 + 1 + 2];
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
            mbar_init(x_empty(s), 4);
            mbar_init(acc_full(s), 1);
        }
        mbar_init(a_ready, 4);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    for (int i = t; i < p.n_bias; i += PF_THREADS) bias_s[i] = p.bias[i];
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t a_u = smem_u32(a_s), w_u = smem_u32(w_s), x_u = smem_u32(x_s);

    if (warp == 4) {
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
                                for (int sub = 0; sub < ntile / 64; ++sub)
                                    pf_tma_2d(dst + (uint32_t)(half * ntile + sub * 64) * (PF_KC * 4), &wmap, 0,
                                              Ly.w_row + (kc * 2 + half) * Ly.N + nt * ntile + sub * 64, w_full(s));
                        }
                }
            }
        }
    } else if (warp == 5) {
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
                            const uint32_t a_hi = a_u + (uint32_t)kc * PF_A_CHUNK_BYTES;
                            const uint32_t a_lo = a_hi + PF_ROWS * PF_KC * 4;
                            const uint32_t b_hi = w_u + (uint32_t)s * PF_STAGE_BYTES;
                            const uint32_t b_lo = b_hi + (uint32_t)ntile * PF_KC * 4;
#pragma unroll
                            for (int ks = 0; ks < PF_KC / 8; ++ks) {
                                const uint32_t ko = ks * 32;
                                const uint64_t dah = umma_desc_k_sw128(a_hi + ko), dal = umma_desc_k_sw128(a_lo + ko);
                                const uint64_t dbh = umma_desc_k_sw128(b_hi + ko), dbl = umma_desc_k_sw128(b_lo + ko);
                                umma_tf32(d, dah, dbh, idesc, (kc | ks) ? 1u : 0u);
                                umma_tf32(d, dal, dbh, idesc, 1u);
                                umma_tf32(d, dah, dbl, idesc, 1u);
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
        const int row = t;  // 0..127 = TMEM lane = point of the superpoint
        uint32_t ci = 0, acc_cnt[2] = {0u, 0u};
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
#pragma unroll
            for (int c16 = 0; c16 < 8; ++c16)
                pf_split_store(a_u, row, c16, make_float4(v[4 * c16], v[4 * c16 + 1], v[4 * c16 + 2], v[4 * c16 + 3]));
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) pf_arrive(a_ready);
            // ---- layers
            for (int l = 0; l < p.n_layers; ++l) {
                const PfLayer& Ly = p.L[l];
                const int ntile = Ly.N < PF_NT ? Ly.N : PF_NT;
                const bool last = l + 1 == p.n_layers;
                for (int nt = 0; nt < Ly.N / ntile; ++nt) {
                    mbar_wait(acc_full(nt), acc_cnt[nt] & 1);
                    ++acc_cnt[nt];
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int cb = 0; cb < ntile / 32; ++cb) {
                        uint32_t r[32];
                        tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(nt * PF_NT + cb * 32), r);
                        const int col0 = nt * ntile + cb * 32;
                        const float* bs = bias_s + Ly.b_off + col0;
                        if (!last) {
                            // next layer's A operand: these 32 columns are exactly K chunk col0/32
                            const uint32_t a_hi = a_u + (uint32_t)(col0 / PF_KC) * PF_A_CHUNK_BYTES;
#pragma unroll
                            for (int c16 = 0; c16 < 8; ++c16) {
                                float4 o;
                                o.x = fmaxf(__uint_as_float(r[4 * c16]) + bs[4 * c16], 0.f);
                                o.y = fmaxf(__uint_as_float(r[4 * c16 + 1]) + bs[4 * c16 + 1], 0.f);
                                o.z = fmaxf(__uint_as_float(r[4 * c16 + 2]) + bs[4 * c16 + 2], 0.f);
                                o.w = fmaxf(__uint_as_float(r[4 * c16 + 3]) + bs[4 * c16 + 3], 0.f);
                                pf_split_store(a_hi, row, c16, o);
                            }
                        } else {
                            // max over the 32 points of this warp: ReLU output is >= 0, so the float order
                            // is the unsigned order of the bit patterns (one redux.sync per column)
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                const float o = fmaxf(__uint_as_float(r[j]) + bs[j], 0.f);
                                const unsigned m = __reduce_max_sync(0xffffffffu, __float_as_uint(o));
                                if (lane == j) pool_s[warp * 256 + col0 + j] = __uint_as_float(m);
                            }
                        }
                    }
                }
                if (!last) {
                    // (all MMAs that read the old A have completed: acc_full of every N-tile was waited for)
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) pf_arrive(a_ready);
                } else {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    asm volatile("bar.sync 1, 128;" ::: "memory");
                    for (int c = t; c < Ly.N; c += 128) {
                        const float m = fmaxf(fmaxf(pool_s[c], pool_s[256 + c]), fmaxf(pool_s[512 + c], pool_s[768 + c]));
                        p.pooled[b * p.ldp + c] = m;
                    }
                    asm volatile("bar.sync 1, 128;" ::: "memory");
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

static int pf_map_2d(CUtensorMap* map, const float* base, uint64_t cols, uint64_t rows, uint32_t box_cols,
                     uint32_t box_rows) {
    PfEncodeFn fn = pf_encode();
    if (!fn) return SPG_E_UNSUPPORTED;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 4};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
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
        if (n != 64 && n != 128 && n != 256) return 0;
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
    int rc = pf_map_2d(&wmap, weight_image, PF_KC, (uint64_t)row, PF_KC, 64);
    if (rc) return rc;
    rc = pf_map_2d(&xmap, clouds, PF_ROWS, (uint64_t)(n_clouds * n_features), PF_ROWS, (uint32_t)n_features);
    if (rc) return rc;
    const int smem = PF_A_BYTES + PF_STAGES * PF_STAGE_BYTES + 2 * PF_X_BYTES + (PF_MAX_BIAS + 4 * 256) * 4;
    cudaError_t e = cudaFuncSetAttribute(pointnet_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    const int64_t grid = n_clouds < kNumSMs ? n_clouds : kNumSMs;
    SPG_LAUNCH(K_POINTNET_FUSED, (cudaStream_t)stream, pointnet_fused_kernel, (unsigned)grid, PF_THREADS, smem, a,
               wmap, xmap);
    return launch_status();
}

}  // extern "C"
