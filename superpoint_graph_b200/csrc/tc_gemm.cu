// tcgen05 (5th-gen tensor core) GEMM for the point-wise PointNet layers:
//
//     C[M,N] = f(A)[M,K] * B[N,K]^T + bias,   f = per-channel affine + ReLU of the producing layer
//
// in fp32-equivalent precision through error-compensated 3xTF32 splitting
// (x = hi + lo with hi = tf32(x), lo = tf32(x - hi);  A*B ~= Ahi*Bhi + Alo*Bhi + Ahi*Blo, all three
// accumulated in the same fp32 TMEM accumulator).  Single-pass TF32 (10-bit mantissa) cannot hold
// the 1e-4 parity bound through five layers + BatchNorm; 3xTF32 is at ~1e-6.
//
// One CTA owns a 128-row tile and the full N (<= 256) -> one UMMA M=128 x N accumulator in TMEM.
// K is streamed in chunks of 32 floats (= one 128-byte swizzle row) through a 2-stage shared
// memory ring.  A is loaded with coalesced 128-bit loads, transformed (affine+ReLU+split) in
// registers and written as K-major SWIZZLE_128B tiles; the weights come as a pre-split,
// pre-swizzled image (tc_pack_weights_kernel) so their load is a straight copy.  A single thread
// issues the tcgen05.mma instructions; tcgen05.commit + mbarriers release stages and publish the
// accumulator.  Epilogue: tcgen05.ld -> registers -> shared staging -> coalesced stores, plus
// the fused batch statistics (count, mean, M2 per 128-row tile) consumed by colstats_merge.
//
// Reference semantics: nn.Conv1d(k=1)+BatchNorm1d+ReLU stacks of learning/pointnet.py:27-37,83-96.
#include "common.cuh"
#include "tc_common.cuh"

namespace spg {

constexpr int TC_BM = 128;      // rows per CTA (UMMA M)
constexpr int TC_KC = 32;       // floats per K chunk (128 B swizzle row)
constexpr int TC_THREADS = 256;
constexpr int TC_A_BYTES = TC_BM * TC_KC * 4;  // 16 KB per (hi|lo) A tile

// Weight image: for every K chunk kc: [hi: N rows x 128 B][lo: N rows x 128 B], each block laid
// out exactly as the shared-memory tile (SWIZZLE_128B).  transpose=0: B[n][k] = W[n*ldw + k];
// transpose=1: B[n][k] = W[k*ldw + n] (the data-gradient GEMM consumes W^T).
__global__ void tc_pack_weights_kernel(const float* __restrict__ W, int64_t ldw, int transpose,
                                       int N, int K, int k_valid, float* __restrict__ img) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)N * K) return;
    const int n = (int)(i / K), k = (int)(i % K);
    float v = 0.f;  // k >= k_valid: zero padding of the reduction dimension
    if (k < k_valid) v = transpose ? W[(int64_t)k * ldw + n] : W[(int64_t)n * ldw + k];
    const uint32_t hi = to_tf32(v);
    const uint32_t lo = to_tf32(v - __uint_as_float(hi));
    const int kc = k / TC_KC, kk = k % TC_KC;
    const int64_t base = (int64_t)kc * 2 * N * TC_KC;
    const int64_t off = (int64_t)(sw128_off(n, kk >> 2) >> 2) + (kk & 3);
    img[base + off] = __uint_as_float(hi);
    img[base + (int64_t)N * TC_KC + off] = __uint_as_float(lo);
}

// Batched variant: one launch packs every weight matrix of a model (forward images and the
// transposed images of the data-gradient GEMMs).  table[j] = {W, ldw, transpose, N, K, k_valid,
// image, first element index of job j}; a thread finds its job by a linear scan (<= 64 jobs).
__global__ void tc_pack_weights_multi_kernel(const long long* __restrict__ table, int n_jobs,
                                             long long total) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int j = 0;
    while (j + 1 < n_jobs && i >= table[(j + 1) * 8 + 7]) ++j;
    const long long* d = table + j * 8;
    const float* W = reinterpret_cast<const float*>(d[0]);
    const long long ldw = d[1];
    const int transpose = (int)d[2], N = (int)d[3], K = (int)d[4], k_valid = (int)d[5];
    float* img = reinterpret_cast<float*>(d[6]);
    const long long e = i - d[7];
    const int n = (int)(e / K), k = (int)(e % K);
    float v = 0.f;
    if (k < k_valid) v = transpose ? W[(long long)k * ldw + n] : W[(long long)n * ldw + k];
    const uint32_t hi = to_tf32(v);
    const uint32_t lo = to_tf32(v - __uint_as_float(hi));
    const int kc = k / TC_KC, kk = k % TC_KC;
    const long long base = (long long)kc * 2 * N * TC_KC;
    const long long off = (long long)(sw128_off(n, kk >> 2) >> 2) + (kk & 3);
    img[base + off] = __uint_as_float(hi);
    img[base + (long long)N * TC_KC + off] = __uint_as_float(lo);
}

struct TcArgs {
    const float* A;
    int64_t lda;
    const float* Wimg;
    const float* bias;
    float* C;
    int64_t ldc;
    int64_t M;
    int K;
    const float *a_scale, *a_shift;
    int a_relu;
    float* stats;  // [row_tiles, N, 3] or null
};

template <int N_>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_gemm_kernel(const TcArgs p) {
    constexpr int W_STAGE_BYTES = 2 * N_ * TC_KC * 4;          // hi + lo
    constexpr int STAGE_BYTES = 2 * TC_A_BYTES + W_STAGE_BYTES;  // A hi, A lo, W hi, W lo
    constexpr int PITCH = N_ + 4;                                // staging row pitch (floats)
    constexpr int TMEM_COLS = N_ < 32 ? 32 : N_;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // manual 1024-byte alignment (SWIZZLE_128B atoms are 1024 B)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[3];  // empty[0], empty[1], done
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int64_t m0 = (int64_t)blockIdx.x * TC_BM;
    const uint32_t bar_empty0 = smem_u32(&bars[0]);
    const uint32_t bar_empty1 = smem_u32(&bars[1]);
    const uint32_t bar_done = smem_u32(&bars[2]);

    if (t == 0) {
        mbar_init(bar_empty0, 1);
        mbar_init(bar_empty1, 1);
        mbar_init(bar_done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    const int nk = p.K / TC_KC;
    const bool pro = p.a_scale || p.a_shift || p.a_relu;
    constexpr uint32_t idesc = umma_idesc_tf32(TC_BM, N_);

    for (int kc = 0; kc < nk; ++kc) {
        const int s = kc & 1;
        uint8_t* stage = smem + (size_t)s * STAGE_BYTES;
        if (kc >= 2) mbar_wait(s ? bar_empty1 : bar_empty0, (uint32_t)(((kc >> 1) - 1) & 1));
        // ---- A chunk: 128 rows x 32 floats, 4 float4 per thread
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = t + TC_THREADS * j;
            const int row = i >> 3, c16 = i & 7;
            const int k = kc * TC_KC + c16 * 4;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m0 + row < p.M) {
                v = __ldg(reinterpret_cast<const float4*>(p.A + (m0 + row) * p.lda + k));
                if (pro) {
                    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.a_scale) sc = __ldg(reinterpret_cast<const float4*>(p.a_scale + k));
                    if (p.a_shift) sh = __ldg(reinterpret_cast<const float4*>(p.a_shift + k));
                    v.x = fmaf(v.x, sc.x, sh.x);
                    v.y = fmaf(v.y, sc.y, sh.y);
                    v.z = fmaf(v.z, sc.z, sh.z);
                    v.w = fmaf(v.w, sc.w, sh.w);
                    if (p.a_relu) {
                        v.x = fmaxf(v.x, 0.f);
                        v.y = fmaxf(v.y, 0.f);
                        v.z = fmaxf(v.z, 0.f);
                        v.w = fmaxf(v.w, 0.f);
                    }
                }
            }
            uint4 hi, lo;
            hi.x = to_tf32(v.x);
            hi.y = to_tf32(v.y);
            hi.z = to_tf32(v.z);
            hi.w = to_tf32(v.w);
            lo.x = to_tf32(v.x - __uint_as_float(hi.x));
            lo.y = to_tf32(v.y - __uint_as_float(hi.y));
            lo.z = to_tf32(v.z - __uint_as_float(hi.z));
            lo.w = to_tf32(v.w - __uint_as_float(hi.w));
            const uint32_t off = sw128_off(row, c16);
            *reinterpret_cast<uint4*>(stage + off) = hi;
            *reinterpret_cast<uint4*>(stage + TC_A_BYTES + off) = lo;
        }
        // ---- W chunk: straight copy of the pre-swizzled image
        {
            const float4* src = reinterpret_cast<const float4*>(p.Wimg + (size_t)kc * (W_STAGE_BYTES / 4));
            float4* dst = reinterpret_cast<float4*>(stage + 2 * TC_A_BYTES);
#pragma unroll
            for (int j = 0; j < W_STAGE_BYTES / 16 / TC_THREADS; ++j)
                dst[t + TC_THREADS * j] = __ldg(src + t + TC_THREADS * j);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(stage), a_lo = a_hi + TC_A_BYTES;
            const uint32_t b_hi = a_hi + 2 * TC_A_BYTES, b_lo = b_hi + W_STAGE_BYTES / 2;
#pragma unroll
            for (int ks = 0; ks < TC_KC / 8; ++ks) {
                const uint32_t ko = ks * 32;  // 8 tf32 = 32 bytes along K inside the swizzle row
                const uint64_t dah = umma_desc_k_sw128(a_hi + ko), dal = umma_desc_k_sw128(a_lo + ko);
                const uint64_t dbh = umma_desc_k_sw128(b_hi + ko), dbl = umma_desc_k_sw128(b_lo + ko);
                umma_tf32(tmem_base, dah, dbh, idesc, (kc | ks) ? 1u : 0u);
                umma_tf32(tmem_base, dal, dbh, idesc, 1u);
                umma_tf32(tmem_base, dah, dbl, idesc, 1u);
            }
            umma_commit(s ? bar_empty1 : bar_empty0);  // frees this stage when the MMAs retire
            if (kc == nk - 1) umma_commit(bar_done);    // accumulator complete
        }
    }

    // ---- epilogue: TMEM -> registers -> shared staging (row-major, pitch N+4)
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    float* stg = reinterpret_cast<float*>(smem);
    {
        const int q = warp & 3;            // TMEM lane quarter this warp may access
        const int half = warp >> 2;        // column half handled by this warp
        const int row = q * 32 + lane;
        constexpr int COLS_PER_WARP = N_ / 2;
#pragma unroll
        for (int cb = 0; cb < COLS_PER_WARP; cb += 32) {
            const int col0 = half * COLS_PER_WARP + cb;
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)col0, r);
            float* drow = stg + row * PITCH + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
                float4 v;
                v.x = __uint_as_float(r[j]) + (p.bias ? __ldg(p.bias + col0 + j) : 0.f);
                v.y = __uint_as_float(r[j + 1]) + (p.bias ? __ldg(p.bias + col0 + j + 1) : 0.f);
                v.z = __uint_as_float(r[j + 2]) + (p.bias ? __ldg(p.bias + col0 + j + 2) : 0.f);
                v.w = __uint_as_float(r[j + 3]) + (p.bias ? __ldg(p.bias + col0 + j + 3) : 0.f);
                *reinterpret_cast<float4*>(drow + j) = v;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)TMEM_COLS)
                     : "memory");
    }
    // ---- coalesced stores
    const int nvalid = (int)min((int64_t)TC_BM, p.M - m0);
    constexpr int F4_PER_ROW = N_ / 4;
    for (int i = t; i < nvalid * F4_PER_ROW; i += TC_THREADS) {
        const int row = i / F4_PER_ROW, c4 = i % F4_PER_ROW;
        const float4 v = *reinterpret_cast<const float4*>(stg + row * PITCH + c4 * 4);
        *reinterpret_cast<float4*>(p.C + (m0 + row) * p.ldc + c4 * 4) = v;
    }
    // ---- fused batch statistics of the tile
    if (p.stats) {
        for (int c = t; c < N_; c += TC_THREADS) {
            float sum = 0.f;
            for (int r = 0; r < nvalid; ++r) sum += stg[r * PITCH + c];
            const float mu = sum / (float)nvalid;
            float m2 = 0.f;
            for (int r = 0; r < nvalid; ++r) {
                const float d = stg[r * PITCH + c] - mu;
                m2 = fmaf(d, d, m2);
            }
            float* o = p.stats + ((int64_t)blockIdx.x * N_ + c) * 3;
            o[0] = (float)nvalid;
            o[1] = mu;
            o[2] = m2;
        }
    }
}

template <int N_>
static int launch_tc(const TcArgs& a, int64_t tiles, cudaStream_t s) {
    constexpr int W_STAGE_BYTES = 2 * N_ * TC_KC * 4;
    constexpr int STAGE_BYTES = 2 * TC_A_BYTES + W_STAGE_BYTES;
    constexpr int STG_BYTES = TC_BM * (N_ + 4) * 4;
    const int smem = (2 * STAGE_BYTES > STG_BYTES ? 2 * STAGE_BYTES : STG_BYTES) + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<N_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    SPG_LAUNCH(K_TC_GEMM, s, tc_gemm_kernel<N_>, (unsigned)tiles, TC_THREADS, smem, a);
    return launch_status();
}

}  // namespace spg

using namespace spg;

extern "C" {

int64_t spg_tc_weight_image_floats(int N, int K) { return (int64_t)2 * N * K; }

static int g_tc_generation = 2;
/* 1 = one CTA per row tile (tc_gemm.cu), 2 = persistent warp-specialised kernel (tc_gemm2.cu). */
int spg_tc_set_generation(int gen) {
    if (gen != 1 && gen != 2) return SPG_E_BADARG;
    g_tc_generation = gen;
    return SPG_OK;
}

/* number of (count, mean, M2) partials per column that spg_tc_gemm writes into stats_ws */
int64_t spg_tc_gemm_stats_partials(int64_t M, int N, int K) {
    const int64_t tiles = M <= 0 ? 1 : ceil_div64(M, TC_BM);
    return (g_tc_generation == 2 && tc_gemm2_handles(N, K)) ? 4 * tiles : tiles;
}

int spg_tc_pack_weights(const float* W, int64_t ldw, int transpose, int N, int K, int k_valid,
                        float* image, spg_stream_t stream) {
    if (!W || !image || N <= 0 || K <= 0 || k_valid <= 0 || k_valid > K) return SPG_E_BADARG;
    if (K % TC_KC != 0 || N % 8 != 0) return SPG_E_UNSUPPORTED;
    const int64_t total = (int64_t)N * K;
    SPG_LAUNCH(K_TC_PACK, (cudaStream_t)stream, tc_pack_weights_kernel,
               (unsigned)ceil_div64(total, 256), 256, 0, W, ldw, transpose, N, K, k_valid, image);
    return launch_status();
}

int spg_tc_pack_weights_multi(const int64_t* table, int n_jobs, int64_t total, spg_stream_t stream) {
    if (!table || n_jobs <= 0 || n_jobs > 64 || total <= 0) return SPG_E_BADARG;
    SPG_LAUNCH(K_TC_PACK, (cudaStream_t)stream, tc_pack_weights_multi_kernel,
               (unsigned)ceil_div64(total, 256), 256, 0, (const long long*)table, n_jobs,
               (long long)total);
    return launch_status();
}

int spg_tc_gemm_supported(int64_t M, int N, int K) {
    return (M > 0 && (N == 64 || N == 128 || N == 256) && K >= TC_KC && K % TC_KC == 0 && K <= 1024) ? 1 : 0;
}

int spg_tc_gemm(const float* A, int64_t lda, const float* weight_image, const float* bias, float* C,
                int64_t ldc, int64_t M, int N, int K, const float* a_scale, const float* a_shift,
                int a_relu, float* stats_ws, spg_stream_t stream) {
    if (M < 0 || !A || !weight_image || !C) return SPG_E_BADARG;
    if (M == 0) return SPG_OK;
    if (!spg_tc_gemm_supported(M, N, K)) return SPG_E_UNSUPPORTED;
    if ((lda & 3) || (ldc & 3) || lda < K || ldc < N) return SPG_E_ALIGN;
    if (((uintptr_t)A | (uintptr_t)C | (uintptr_t)weight_image | (uintptr_t)a_scale |
         (uintptr_t)a_shift) & 15)
        return SPG_E_ALIGN;
    TcArgs a;
    a.A = A; a.lda = lda; a.Wimg = weight_image; a.bias = bias; a.C = C; a.ldc = ldc; a.M = M;
    a.K = K; a.a_scale = a_scale; a.a_shift = a_shift; a.a_relu = a_relu; a.stats = stats_ws;
    const int64_t tiles = ceil_div64(M, TC_BM);
    if (tiles > 2147483647ll) return SPG_E_UNSUPPORTED;
    cudaStream_t s = (cudaStream_t)stream;
    if (g_tc_generation == 2) {
        bool handled = false;
        const int rc = tc_gemm2_try(A, lda, weight_image, bias, C, ldc, M, N, K, a_scale, a_shift,
                                    a_relu, stats_ws, s, &handled);
        if (handled) return rc;
    }
    if (N == 64) return launch_tc<64>(a, tiles, s);
    if (N == 128) return launch_tc<128>(a, tiles, s);
    return launch_tc<256>(a, tiles, s);
}

}  // extern "C"
