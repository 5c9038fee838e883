// tcgen05 weight-gradient GEMM of the point-wise layers:
//
//     dW[Co, Ci] = sum_m dY[m, Co] * f(P)[m, Ci]        (m runs over all Nv*L points)
//
// Both operands are read exactly as they lie in HBM (point-major rows, channels contiguous), i.e.
// the reduction dimension is the OUTER one: for the tensor core this is an "MN-major x MN-major"
// product, D[Co,Ci] += A[Co,k] * B[Ci,k]^T with k = points.  A CTA owns a contiguous slab of points,
// streams it in chunks of 32 points through a 2-stage shared-memory ring (MN-major SWIZZLE_128B
// SW128_32B atoms: 32 channels x 4 points = 512 B), splits every value into tf32 hi/lo on the fly (3xTF32,
// fp32-equivalent), keeps the full [Co,Ci] accumulator in TMEM (Co/128 tiles x Ci columns) for
// its whole lifetime and finally writes one partial per CTA; gemm_splitk_reduce adds the partials
// in a fixed order (deterministic).  f = affine + ReLU of the layer that produced P (fused).
//
// Reference semantics: the weight gradient of nn.Conv1d(k=1) (learning/pointnet.py:29,85) as
// autograd computes it; the reference materialises ReLU(BN(P)) and runs cuDNN/cuBLAS on it.
#include "common.cuh"
#include "tc_common.cuh"

namespace spg {

constexpr int DW_THREADS = 256;
constexpr int DW_PTS = 32;  // points per chunk (4 UMMA K-steps of 8)

struct DwArgs {
    const float* dY;
    int64_t lddy;
    const float* P;
    int64_t ldp;
    const float *p_scale, *p_shift;
    int p_relu;
    float* partial;  // [gridDim.x, CO, CI]
    int64_t M;
    int64_t pts_per_cta;  // multiple of DW_PTS
};

// byte offset of (point pt in [0,32), 16-byte channel chunk c4) in an MN-major SW128_32B tile with
// MB 32-channel blocks: 512-byte atoms (32 channels x 4 points) ordered [point group][channel block];
// inside an atom the 32-byte chunk index is XORed with the point row (Swizzle<2,5,2>).
template <int MB>
__device__ __forceinline__ uint32_t mn_off(int pt, int c4) {
    const int kg = pt >> 2, kr = pt & 3, mb = c4 >> 3, j = c4 & 7;
    return (uint32_t)((kg * MB + mb) * 512 + kr * 128 + (((j >> 1) ^ kr) << 5) + ((j & 1) << 4));
}

template <int CO, int CI>
__global__ void __launch_bounds__(DW_THREADS, 1) tc_dw_kernel(const DwArgs p) {
    constexpr int CO_PAD = CO < 128 ? 128 : CO;  // the accumulator tile has 128 rows (UMMA M)
    constexpr int MB_A = CO_PAD / 32, MB_B = CI / 32;
    constexpr int A_BYTES = CO_PAD * DW_PTS * 4;  // one of hi|lo
    constexpr int B_BYTES = CI * DW_PTS * 4;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    constexpr int MT = CO_PAD / 128;              // 128-row accumulator tiles
    constexpr int TMEM_COLS_RAW = MT * CI;
    constexpr int TMEM_COLS = TMEM_COLS_RAW <= 32 ? 32 : TMEM_COLS_RAW <= 64 ? 64 : TMEM_COLS_RAW <= 128 ? 128 : TMEM_COLS_RAW <= 256 ? 256 : 512;
    static_assert(CO % 64 == 0 && CI % 32 == 0 && CI <= 256 && TMEM_COLS_RAW <= 512, "shape");
    constexpr int A_F4 = CO * DW_PTS / 4 / DW_THREADS;  // float4 per thread per chunk
    constexpr int B_F4 = CI * DW_PTS / 4 / DW_THREADS;

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ __align__(8) uint64_t bars[3];
    __shared__ uint32_t tmem_base_s;

    const int t = threadIdx.x;
    const int warp = t >> 5, lane = t & 31;
    const uint32_t bar_empty0 = smem_u32(&bars[0]), bar_empty1 = smem_u32(&bars[1]);
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
    if (CO < CO_PAD) {
        // channels CO..127 of the A tiles are never written by the loads: zero them once
        for (int i = t; i < 2 * STAGE_BYTES / 16; i += DW_THREADS)
            reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncthreads();
    }
    SPG_PDL_ENTRY();  // set-up above overlaps the previous kernel of the stream; global memory only below

    const int64_t m_beg = (int64_t)blockIdx.x * p.pts_per_cta;
    const int64_t m_end = min(p.M, m_beg + p.pts_per_cta);
    const int nchunks = m_beg < m_end ? (int)((m_end - m_beg + DW_PTS - 1) / DW_PTS) : 0;
    constexpr uint32_t idesc = umma_idesc_tf32_major(128, CI, 1, 1);
    const bool pro = p.p_scale || p.p_shift || p.p_relu;

    float4 ra[A_F4], rb[B_F4];
    auto load_chunk = [&](int ch) {
        const int64_t mrow = m_beg + (int64_t)ch * DW_PTS;
#pragma unroll
        for (int j = 0; j < A_F4; ++j) {
            const int i = t + DW_THREADS * j;
            const int pt = i / (CO / 4), c4 = i % (CO / 4);
            ra[j] = (mrow + pt < m_end)
                        ? __ldg(reinterpret_cast<const float4*>(p.dY + (mrow + pt) * p.lddy + c4 * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < B_F4; ++j) {
            const int i = t + DW_THREADS * j;
            const int pt = i / (CI / 4), c4 = i % (CI / 4);
            rb[j] = (mrow + pt < m_end)
                        ? __ldg(reinterpret_cast<const float4*>(p.P + (mrow + pt) * p.ldp + c4 * 4))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto split_store = [&](uint8_t* hi_base, uint8_t* lo_base, uint32_t off, float4 v) {
        uint4 hi, lo;
        hi.x = to_tf32(v.x);
        hi.y = to_tf32(v.y);
        hi.z = to_tf32(v.z);
        hi.w = to_tf32(v.w);
        lo.x = to_tf32(v.x - __uint_as_float(hi.x));
        lo.y = to_tf32(v.y - __uint_as_float(hi.y));
        lo.z = to_tf32(v.z - __uint_as_float(hi.z));
        lo.w = to_tf32(v.w - __uint_as_float(hi.w));
        *reinterpret_cast<uint4*>(hi_base + off) = hi;
        *reinterpret_cast<uint4*>(lo_base + off) = lo;
    };

    if (nchunks > 0) load_chunk(0);
    for (int ch = 0; ch < nchunks; ++ch) {
        const int s = ch & 1;
        uint8_t* stage = smem + (size_t)s * STAGE_BYTES;
        uint8_t* a_hi = stage;
        uint8_t* a_lo = stage + A_BYTES;
        uint8_t* b_hi = stage + 2 * A_BYTES;
        uint8_t* b_lo = b_hi + B_BYTES;
        if (ch >= 2) mbar_wait(s ? bar_empty1 : bar_empty0, (uint32_t)(((ch >> 1) - 1) & 1));
        const int64_t mrow = m_beg + (int64_t)ch * DW_PTS;
#pragma unroll
        for (int j = 0; j < A_F4; ++j) {
            const int i = t + DW_THREADS * j;
            const int pt = i / (CO / 4), c4 = i % (CO / 4);
            split_store(a_hi, a_lo, mn_off<MB_A>(pt, c4), ra[j]);
        }
#pragma unroll
        for (int j = 0; j < B_F4; ++j) {
            const int i = t + DW_THREADS * j;
            const int pt = i / (CI / 4), c4 = i % (CI / 4);
            float4 v = rb[j];
            if (pro && (mrow + pt < m_end)) {
                float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.p_scale) sc = __ldg(reinterpret_cast<const float4*>(p.p_scale + c4 * 4));
                if (p.p_shift) sh = __ldg(reinterpret_cast<const float4*>(p.p_shift + c4 * 4));
                v.x = fmaf(v.x, sc.x, sh.x);
                v.y = fmaf(v.y, sc.y, sh.y);
                v.z = fmaf(v.z, sc.z, sh.z);
                v.w = fmaf(v.w, sc.w, sh.w);
                if (p.p_relu) {
                    v.x = fmaxf(v.x, 0.f);
                    v.y = fmaxf(v.y, 0.f);
                    v.z = fmaxf(v.z, 0.f);
                    v.w = fmaxf(v.w, 0.f);
                }
            }
            split_store(b_hi, b_lo, mn_off<MB_B>(pt, c4), v);
        }
        if (ch + 1 < nchunks) load_chunk(ch + 1);  // in flight while the tensor core works
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (t == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo);
            const uint32_t bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
            for (int kb = 0; kb < DW_PTS / 8; ++kb) {
                // one K-step = 8 points = two 4-point groups (sbo apart)
                const uint64_t dbh = umma_desc_mn_sw128_32b(bh + 2 * kb * MB_B * 512, 512, MB_B * 512);
                const uint64_t dbl = umma_desc_mn_sw128_32b(bl + 2 * kb * MB_B * 512, 512, MB_B * 512);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const uint32_t aoff = (2 * kb * MB_A + mt * 4) * 512;
                    const uint64_t dah = umma_desc_mn_sw128_32b(ah + aoff, 512, MB_A * 512);
                    const uint64_t dal = umma_desc_mn_sw128_32b(al + aoff, 512, MB_A * 512);
                    const uint32_t d = tmem_base + mt * CI;
                    umma_tf32(d, dah, dbh, idesc, (ch | kb) ? 1u : 0u);
                    umma_tf32(d, dal, dbh, idesc, 1u);
                    umma_tf32(d, dah, dbl, idesc, 1u);
                }
            }
            umma_commit(s ? bar_empty1 : bar_empty0);
            if (ch == nchunks - 1) umma_commit(bar_done);
        }
    }

    // ---- epilogue: accumulator -> this CTA's partial [CO, CI]
    float* out = p.partial + (int64_t)blockIdx.x * CO * CI;
    if (nchunks > 0) {
        mbar_wait(bar_done, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    {
        const int q = warp & 3, half = warp >> 2;
        constexpr int NBLK = CI / 32;  // 32-column blocks per accumulator tile
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int row = mt * 128 + q * 32 + lane;
            if (row >= CO) continue;  // padded accumulator rows
            for (int cb = half; cb < NBLK; cb += 2) {
                float4* dst = reinterpret_cast<float4*>(out + (int64_t)row * CI + cb * 32);
                if (nchunks > 0) {
                    uint32_t r[32];
                    tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * CI + cb * 32), r);
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        dst[j] = make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                             __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) dst[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
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
}

template <int CO, int CI>
static int launch_dw(const DwArgs& a, int ctas, cudaStream_t s) {
    constexpr int CO_PAD = CO < 128 ? 128 : CO;
    constexpr int STAGE_BYTES = 2 * CO_PAD * DW_PTS * 4 + 2 * CI * DW_PTS * 4;
    const int smem = 2 * STAGE_BYTES + 1024;
    cudaError_t e = cudaFuncSetAttribute(tc_dw_kernel<CO, CI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return (int)e;
    SPG_LAUNCH(K_TC_DW, s, (tc_dw_kernel<CO, CI>), (unsigned)ctas, DW_THREADS, smem, a);
    return launch_status();
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_tc_dw_supported(int64_t M, int co, int ci) {
    const bool co_ok = (co == 64 || co == 128 || co == 256);
    const bool ci_ok = (ci == 32 || ci == 64 || ci == 128);
    return (M >= DW_PTS && co_ok && ci_ok) ? 1 : 0;
}

int spg_tc_dw_ctas(int64_t M) {
    int64_t chunks = ceil_div64(M, DW_PTS);
    int64_t ctas = chunks < kNumSMs ? chunks : kNumSMs;
    return (int)(ctas < 1 ? 1 : ctas);
}

int spg_tc_dw(const float* dY, int64_t lddy, const float* P, int64_t ldp, const float* p_scale,
              const float* p_shift, int p_relu, float* dW, float* workspace, int64_t M, int co, int ci,
              spg_stream_t stream) {
    if (!dY || !P || !dW || !workspace || M <= 0) return SPG_E_BADARG;
    if (!spg_tc_dw_supported(M, co, ci)) return SPG_E_UNSUPPORTED;
    if ((lddy & 3) || (ldp & 3) || lddy < co || ldp < ci) return SPG_E_ALIGN;
    if (((uintptr_t)dY | (uintptr_t)P | (uintptr_t)workspace | (uintptr_t)p_scale | (uintptr_t)p_shift) & 15)
        return SPG_E_ALIGN;
    const int ctas = spg_tc_dw_ctas(M);
    DwArgs a;
    a.dY = dY; a.lddy = lddy; a.P = P; a.ldp = ldp; a.p_scale = p_scale; a.p_shift = p_shift;
    a.p_relu = p_relu; a.partial = workspace; a.M = M;
    cudaStream_t s = (cudaStream_t)stream;
    a.pts_per_cta = ceil_div64(ceil_div64(M, ctas), DW_PTS) * DW_PTS;
    int rc;
#define SPG_DW_CASE(CO_, CI_) \
    if (co == CO_ && ci == CI_) rc = launch_dw<CO_, CI_>(a, ctas, s); else
    SPG_DW_CASE(64, 32) SPG_DW_CASE(64, 64) SPG_DW_CASE(64, 128) SPG_DW_CASE(128, 32)
    SPG_DW_CASE(128, 64) SPG_DW_CASE(128, 128) SPG_DW_CASE(256, 32) SPG_DW_CASE(256, 64)
    SPG_DW_CASE(256, 128) rc = SPG_E_UNSUPPORTED;
#undef SPG_DW_CASE
    if (rc) return rc;
    return spg_splitk_reduce(workspace, ctas, co, ci, nullptr, dW, ci, stream);
}

}  // extern "C"
