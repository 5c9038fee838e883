// tcgen05 3xTF32 GEMM, second generation: persistent, warp-specialised, weights resident in SMEM.
//
//     C[M,N] = f(A)[M,K] * B[N,K]^T + bias     (same contract as tc_gemm.cu / spg_tc_gemm)
//
// Measured problem of the first-generation kernel (profiles/README.md): every 128-row tile re-streamed
// the whole weight image from L2 (64 KB per K-chunk at N=256) and nothing overlapped across tiles.
// Here a CTA (one per SM) owns an N-slice of <= 128 output channels whose pre-split weight image
// (hi+lo, <= 128 KB) is loaded into shared memory ONCE, and then walks over row tiles:
//
//   warps 0-3   epilogue : tcgen05.ld accumulator -> +bias -> 128-bit stores, fused batch statistics
//                          (per 32-row group: pivoted sums via a shuffle transpose-reduce)
//   warp  4     MMA      : one thread issues 12 tcgen05.mma per K-chunk, tcgen05.commit -> mbarriers
//   warps 5-12  producer : coalesced 128-bit loads of A (next chunk prefetched in registers), fused
//                          affine+ReLU, tf32 hi/lo split, SWIZZLE_128B K-major tiles, 2-stage ring
//
// Two TMEM accumulators (2 x NS columns) let the epilogue of tile i overlap the MMAs of tile i+1.
// mbarriers: full[s] (256 producer arrivals), empty[s] (commit), accfull[a] (commit),
// accempty[a] (128 epilogue arrivals).
#include <stdlib.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace spg {

constexpr int T2_BM = 128;
constexpr int T2_KC = 32;
constexpr int T2_MAX_STAGES = 4;  // the A ring gets as many stages as fit next to the resident weights
constexpr int T2_EPI_WARPS = 4, T2_PROD_WARPS = 8;  // (8 epilogue warps force 96 regs/thread: measured slower)
constexpr int T2_THREADS = (T2_EPI_WARPS + 1 + T2_PROD_WARPS) * 32;  // 416
constexpr int T2_A_BYTES = T2_BM * T2_KC * 4;                        // 16 KB (hi or lo)
constexpr int T2_STAGE_BYTES = 2 * T2_A_BYTES;
constexpr int T2_EPI_PITCH = 36;  // floats per row of an epilogue transpose tile (16-byte aligned rows)

struct Tc2Args {
    const float* A;
    int64_t lda;
    const float* Wimg;  // full image [K/32][hi|lo][N][32] (tc_pack_weights)
    const float* bias;
    float* C;
    int64_t ldc;
    int64_t M;
    int N, K;
    const float *a_scale, *a_shift;
    int a_relu;
    float* stats;  // [4*tiles, N, 3] or null
    int nstages;   // A-ring depth (2..4)
    int dbg;       // experiment switches (SPG_TC_DBG): 1 no epilogue stores, 2 no MMA, 4 no loads, 8 no STS
};

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(T2_THREADS, 1) tc_gemm2_kernel(const Tc2Args p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // the dynamic segment starts 1024-byte aligned (declared alignment; static data is padded up to
    // it): SWIZZLE_128B atoms need that, and no spare bytes are reserved for a manual round-up
    uint8_t* smem = smem_raw;
    if ((smem_u32(smem_raw) & 1023u) != 0u) __trap();
    __shared__ __align__(8) uint64_t bars[2 * T2_MAX_STAGES + 4];
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const int nk = p.K / T2_KC;
    const int n0 = blockIdx.y * NS;  // first output channel of this CTA's slice
    const int64_t tiles = (p.M + T2_BM - 1) / T2_BM;
    const int nst = p.nstages;
    uint8_t* wres = smem + (size_t)nst * T2_STAGE_BYTES;  // resident weights: [nk][hi|lo][NS][128 B]

    const uint32_t bars_u32 = smem_u32(&bars[0]);
    auto bar_full = [&](int s) { return bars_u32 + 8u * (uint32_t)s; };
    auto bar_empty = [&](int s) { return bars_u32 + 8u * (uint32_t)(T2_MAX_STAGES + s); };
    auto bar_accfull = [&](uint32_t a) { return bars_u32 + 8u * (2 * T2_MAX_STAGES + a); };
    auto bar_accempty = [&](uint32_t a) { return bars_u32 + 8u * (2 * T2_MAX_STAGES + 2 + a); };

    if (t == 0) {
#pragma unroll
        for (int s = 0; s < T2_MAX_STAGES; ++s) {
            mbar_init(bar_full(s), T2_PROD_WARPS);  // one elected arrival per producer warp
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_accfull(0), 1);
        mbar_init(bar_accfull(1), 1);
        mbar_init(bar_accempty(0), T2_EPI_WARPS);  // one elected arrival per epilogue warp
        mbar_init(bar_accempty(1), T2_EPI_WARPS);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                         smem_u32(&tmem_base_s)),
                     "r"((uint32_t)(2 * NS))
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Producer threads put their first A chunks in flight before anything else: the global-load
    // latency then overlaps the resident-weight load and the CTA-wide barrier below.
    // Register-level prefetch ring: PF chunks of A are in flight per thread (the global-load
    // latency, ~2 us under load, is far longer than one chunk's transform + MMA).
    constexpr int PF = 4;
    float4 q[PF][4];
    const int pt = t - (T2_EPI_WARPS + 1) * 32;  // producer thread id 0..255 (negative: other roles)
    auto load = [&](int64_t tile, int kc, float4 (&dst)[4]) {
        const int64_t m0 = tile * T2_BM;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = pt + 256 * j;
            const int row = i >> 3, c16 = i & 7;
            dst[j] = (!(p.dbg & 4) && tile < tiles && m0 + row < p.M)
                         ? __ldg(reinterpret_cast<const float4*>(p.A + (m0 + row) * p.lda +
                                                                 kc * T2_KC + c16 * 4))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto advance = [&](int64_t& tl, int& k) {
        if (++k == nk) {
            k = 0;
            tl += gridDim.x;
        }
    };
    // load cursor (runs PF-1 items ahead of the consume cursor)
    int64_t ltile = blockIdx.x;
    int lkc = 0;
    if (pt >= 0) {
#pragma unroll
        for (int d = 0; d < PF - 1; ++d) {
            load(ltile, lkc, q[d]);
            advance(ltile, lkc);
        }
    }
    // resident weight slice: rows [n0, n0+NS) of every (chunk, hi|lo) block of the image
    {
        const int f4_per_block = NS * T2_KC / 4;  // float4 per (chunk, half) block of the slice
        const int total = nk * 2 * f4_per_block;
        for (int i = t; i < total; i += T2_THREADS) {
            const int blk = i / f4_per_block, r = i % f4_per_block;
            const float4* src = reinterpret_cast<const float4*>(
                p.Wimg + ((int64_t)blk * p.N + n0) * T2_KC);
            reinterpret_cast<float4*>(wres)[(int64_t)blk * f4_per_block + r] = __ldg(src + r);
        }
    }
    // per-channel vectors of the fused prologue / epilogue, once per CTA
    float* sc_s = reinterpret_cast<float*>(wres + (size_t)nk * 2 * NS * T2_KC * 4);
    float* sh_s = sc_s + p.K;
    float* bias_s = sh_s + p.K;
    // per-epilogue-warp transpose tiles [32 rows][36]: accumulator rows (one per lane) are turned
    // into full 128-byte lines before they go to global memory
    float* epi_s = bias_s + NS;
    for (int i = t; i < p.K; i += T2_THREADS) {
        sc_s[i] = p.a_scale ? p.a_scale[i] : 1.f;
        sh_s[i] = p.a_shift ? p.a_shift[i] : 0.f;
    }
    for (int i = t; i < NS; i += T2_THREADS) bias_s[i] = p.bias ? p.bias[n0 + i] : 0.f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_s;

    if (warp >= T2_EPI_WARPS + 1) {
        // ================================ producers ================================
        const bool pro = p.a_scale || p.a_shift || p.a_relu;
        int64_t tile = blockIdx.x;
        int kc = 0;
        uint32_t it = 0;
        const int c16 = pt & 7;  // 16-byte chunk of the 128-byte K row (same for all 4 rows)
        uint32_t soff[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) soff[j] = sw128_off((pt >> 3) + 32 * j, c16);
        const uint32_t smem_u = smem_u32(smem);
        auto st_shared4 = [](uint32_t addr, uint4 v) {
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y),
                         "r"(v.z), "r"(v.w)
                         : "memory");
        };
        auto consume = [&](float4 (&cur)[4], float4 (&far)[4]) {
            const int s = it % nst;
            const uint32_t use = it / nst;
            load(ltile, lkc, far);  // item it+PF-1
            advance(ltile, lkc);
            const float4 sc = *reinterpret_cast<const float4*>(sc_s + kc * T2_KC + c16 * 4);
            const float4 sh = *reinterpret_cast<const float4*>(sh_s + kc * T2_KC + c16 * 4);
            if (use > 0) mbar_wait(bar_empty(s), (use - 1) & 1);
            const uint32_t stage = smem_u + (uint32_t)s * T2_STAGE_BYTES;
            const int64_t m0 = tile * T2_BM;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 v = cur[j];
                if (pro && (m0 + (pt >> 3) + 32 * j < p.M)) {
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
                uint4 hi, lo;
                hi.x = to_tf32(v.x);
                hi.y = to_tf32(v.y);
                hi.z = to_tf32(v.z);
                hi.w = to_tf32(v.w);
                lo.x = to_tf32(v.x - __uint_as_float(hi.x));
                lo.y = to_tf32(v.y - __uint_as_float(hi.y));
                lo.z = to_tf32(v.z - __uint_as_float(hi.z));
                lo.w = to_tf32(v.w - __uint_as_float(hi.w));
                if (!(p.dbg & 8)) {
                    st_shared4(stage + soff[j], hi);
                    st_shared4(stage + T2_A_BYTES + soff[j], lo);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full(s));
            advance(tile, kc);
            ++it;
        };
        while (tile < tiles) {
            consume(q[0], q[3]);
            if (tile >= tiles) break;
            consume(q[1], q[0]);
            if (tile >= tiles) break;
            consume(q[2], q[1]);
            if (tile >= tiles) break;
            consume(q[3], q[2]);
        }
    } else if (warp == T2_EPI_WARPS) {
        // ================================ MMA issuer ================================
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_tf32(T2_BM, NS);
            const uint32_t wres_u32 = smem_u32(wres);
            uint32_t it = 0, tcount = 0;
            for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
                const uint32_t a = tcount & 1;
                if (tcount >= 2) mbar_wait(bar_accempty(a), ((tcount >> 1) - 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + a * NS;
                for (int kc = 0; kc < nk; ++kc, ++it) {
                    const int s = it % nst;
                    mbar_wait(bar_full(s), (it / nst) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * T2_STAGE_BYTES);
                    const uint32_t a_lo = a_hi + T2_A_BYTES;
                    const uint32_t b_hi = wres_u32 + (uint32_t)(kc * 2) * (NS * T2_KC * 4);
                    const uint32_t b_lo = b_hi + NS * T2_KC * 4;
#pragma unroll
                    for (int ks = 0; ks < T2_KC / 8; ++ks) {
                        const uint32_t ko = ks * 32;
                        const uint64_t dah = umma_desc_k_sw128(a_hi + ko), dal = umma_desc_k_sw128(a_lo + ko);
                        const uint64_t dbh = umma_desc_k_sw128(b_hi + ko), dbl = umma_desc_k_sw128(b_lo + ko);
                        if (!(p.dbg & 2)) {
                            umma_tf32(d, dah, dbh, idesc, (kc | ks) ? 1u : 0u);
                            umma_tf32(d, dal, dbh, idesc, 1u);
                            umma_tf32(d, dah, dbl, idesc, 1u);
                        }
                    }
                    umma_commit(bar_empty(s));
                }
                umma_commit(bar_accfull(a));
            }
        }
    } else {
        // ================================ epilogue ================================
        const int w = warp & 3;      // TMEM lane quarter this warp may access
        const int cb0 = warp >> 2;   // the two warps of a quarter take alternate 32-column blocks
        uint32_t tcount = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++tcount) {
            const uint32_t a = tcount & 1;
            mbar_wait(bar_accfull(a), (tcount >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t row0 = tile * T2_BM + w * 32;
            const float nvalid = (float)max((int64_t)0, min((int64_t)32, p.M - row0));
#pragma unroll 1
            for (int cb = cb0; cb < NS / 32; cb += T2_EPI_WARPS / 4) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(a * NS + cb * 32), r);
                const int col0 = n0 + cb * 32;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    v[j] = __uint_as_float(r[j]) + bias_s[cb * 32 + j];
                if (!(p.dbg & 1)) {
                    // lane = row -> shared tile -> (4 rows x 128 B) per store instruction
                    float* tile_s = epi_s + warp * (32 * T2_EPI_PITCH);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(tile_s + lane * T2_EPI_PITCH + 4 * j) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    const int rr = lane >> 3, c4 = lane & 7;
#pragma unroll
                    for (int itr = 0; itr < 8; ++itr) {
                        const int r = itr * 4 + rr;
                        const float4 q = *reinterpret_cast<const float4*>(tile_s + r * T2_EPI_PITCH + 4 * c4);
                        if (row0 + r < p.M)
                            *reinterpret_cast<float4*>(p.C + (row0 + r) * p.ldc + col0 + 4 * c4) = q;
                    }
                    __syncwarp();
                }
                if (p.stats && !(p.dbg & 1)) {
                    // per-column sums over the 32 rows of this warp, read back from the transpose
                    // tile (lane = column: conflict-free), pivoted on the group's first row
                    const float* tile_c = epi_s + warp * (32 * T2_EPI_PITCH) + lane;
                    // (the tile still holds this block: the stores above only read it)
                    const float pivot = tile_c[0];
                    float d1 = 0.f, d2 = 0.f;
                    const int nv = (int)nvalid;
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const float d = tile_c[r * T2_EPI_PITCH] - pivot;
                        if (r < nv) {
                            d1 += d;
                            d2 = fmaf(d, d, d2);
                        }
                    }
                    float* o = p.stats + (((int64_t)tile * 4 + w) * p.N + col0 + lane) * 3;
                    if (nvalid > 0.f) {
                        o[0] = nvalid;
                        o[1] = pivot + d1 / nvalid;
                        o[2] = fmaxf(d2 - d1 * d1 / nvalid, 0.f);
                    } else {
                        o[0] = 0.f;
                        o[1] = 0.f;
                        o[2] = 0.f;
                    }
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_accempty(a));
        }
    }

    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)(2 * NS))
                     : "memory");
    }
}

template <int NS>
static int launch_tc2(const Tc2Args& a, cudaStream_t s) {
    const int slices = a.N / NS;
    const int64_t tiles = ceil_div64(a.M, T2_BM);
    int64_t gx = kNumSMs / slices;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    const int fixed = (a.K / T2_KC) * 2 * NS * T2_KC * 4 + (2 * a.K + NS) * 4 +
                      T2_EPI_WARPS * 32 * T2_EPI_PITCH * 4;
    int nst = (232448 - 1024 - fixed) / T2_STAGE_BYTES;  // 227 KB per CTA minus 1 KB static
    if (nst > T2_MAX_STAGES) nst = T2_MAX_STAGES;
    if (nst < 2) return SPG_E_UNSUPPORTED;
    Tc2Args a2 = a;
    a2.nstages = nst;
    const int smem = nst * T2_STAGE_BYTES + fixed;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm2_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid((unsigned)gx, (unsigned)slices);
    SPG_LAUNCH(K_TC_GEMM, s, tc_gemm2_kernel<NS>, grid, T2_THREADS, smem, a2);
    return launch_status();
}

// slice width: the resident image (2*NS*K*4 bytes) must fit next to the A ring in 227 KB
static inline int pick_ns(int N, int K) {
    if (N % 128 == 0 && 128 * K <= 16384) return 128;
    if (N % 64 == 0 && 64 * K <= 16384) return 64;
    return 0;
}

int tc_gemm2_try(const float* A, int64_t lda, const float* weight_image, const float* bias, float* C,
                 int64_t ldc, int64_t M, int N, int K, const float* a_scale, const float* a_shift,
                 int a_relu, float* stats_ws, cudaStream_t s, bool* handled) {
    const int ns = pick_ns(N, K);
    *handled = ns != 0;
    if (!ns) return SPG_OK;
    Tc2Args a;
    a.A = A; a.lda = lda; a.Wimg = weight_image; a.bias = bias; a.C = C; a.ldc = ldc; a.M = M;
    a.N = N; a.K = K; a.a_scale = a_scale; a.a_shift = a_shift; a.a_relu = a_relu; a.stats = stats_ws;
    {
        static int dbg = -1;
        if (dbg < 0) {
            const char* e = getenv("SPG_TC_DBG");
            dbg = e ? atoi(e) : 0;
        }
        a.dbg = dbg;
    }
    return ns == 128 ? launch_tc2<128>(a, s) : launch_tc2<64>(a, s);
}

bool tc_gemm2_handles(int N, int K) { return pick_ns(N, K) != 0; }

}  // namespace spg
