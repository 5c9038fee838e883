// Fused GRUCellEx (GRU cell + input gate + affine-free layer norm of both gate
// pre-activations), forward and backward, one kernel each.
//
// Reference semantics: learning/modules.py:205-251.  There the cell is ~20 torch
// ops per call (3 GEMMs, 2 InstanceNorm1d, chunk/sigmoid/tanh/elementwise); here the
// three weight matrices (7*H*H floats, 28 KB at H=32) live in shared memory,
// a warp owns RW rows at a time and every intermediate stays on chip.
// 14 kFLOP and 384 B per row: latency/L2-bound, so no tensor cores.
#include "common.cuh"

namespace spg {

constexpr int kGruWarps = 8;  // warps per block

// shared-memory layout (floats):
//   Wig_t [H][H+1]    Wig_t[k*(H+1)+c]   = ig_weight[c][k]
//   Wih_t [H][3H+1]   Wih_t[k*(3H+1)+j]  = weight_ih[j][k]
//   Whh_t [H][3H+1]
//   per warp scratch: hrow[RW][H], xrow[RW][H], srow[RW][H], gi[RW][3H], gh[RW][3H]
__host__ __device__ inline int gru_weight_floats(int H) { return H * (H + 1) + 2 * H * (3 * H + 1); }
__host__ __device__ inline int gru_scratch_floats(int H, int rw) { return rw * (3 * H + 6 * H); }

__device__ __forceinline__ void gru_load_weights(float* sm, const float* __restrict__ w_ih,
                                                 const float* __restrict__ w_hh,
                                                 const float* __restrict__ w_ig, int H,
                                                 int ingate) {
    float* Wig_t = sm;
    float* Wih_t = Wig_t + H * (H + 1);
    float* Whh_t = Wih_t + H * (3 * H + 1);
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < 3 * H * H; i += nt) {
        const int j = i / H, k = i % H;
        Wih_t[k * (3 * H + 1) + j] = w_ih[i];
        Whh_t[k * (3 * H + 1) + j] = w_hh[i];
    }
    if (ingate) {
        for (int i = tid; i < H * H; i += nt) {
            const int c = i / H, k = i % H;
            Wig_t[k * (H + 1) + c] = w_ig[i];
        }
    }
}

// Recomputes everything up to the normalised gate inputs for RW rows.
// On return (per row i): hrow = h, xrow = gated input x', srow = sigmoid(q) (or 1),
// gi/gh = raw (pre-norm) gate inputs, stats = {mean_i, rstd_i, mean_h, rstd_h}.
template <int kRW>
__device__ __forceinline__ void gru_rows_forward(const float* sm, float* scratch, int H, int flags,
                                                 const float* __restrict__ x,
                                                 const float* __restrict__ h,
                                                 const float* __restrict__ b_ig, int64_t row0,
                                                 int64_t n_rows, int lane, float stats[kRW][4]) {
    const float* Wig_t = sm;
    const float* Wih_t = Wig_t + H * (H + 1);
    const float* Whh_t = Wih_t + H * (3 * H + 1);
    float* hrow = scratch;
    float* xrow = hrow + kRW * H;
    float* srow = xrow + kRW * H;
    float* gi = srow + kRW * H;
    float* gh = gi + kRW * 3 * H;
    const int H3 = 3 * H;

    for (int i = 0; i < kRW; ++i) {
        const int64_t row = row0 + i;
        for (int c = lane; c < H; c += 32) {
            hrow[i * H + c] = row < n_rows ? h[row * H + c] : 0.f;
            xrow[i * H + c] = row < n_rows ? x[row * H + c] : 0.f;
        }
    }
    __syncwarp();
    if (flags & SPG_GRU_INGATE) {
        for (int c = lane; c < H; c += 32) {
            float acc[kRW];
#pragma unroll
            for (int i = 0; i < kRW; ++i) acc[i] = b_ig[c];
            for (int k = 0; k < H; ++k) {
                const float wv = Wig_t[k * (H + 1) + c];
#pragma unroll
                for (int i = 0; i < kRW; ++i) acc[i] = fmaf(wv, hrow[i * H + k], acc[i]);
            }
#pragma unroll
            for (int i = 0; i < kRW; ++i) {
                const float sg = sigmoidf_(acc[i]);
                srow[i * H + c] = sg;
                xrow[i * H + c] *= sg;  // only this lane touches xrow[.][c]
            }
        }
    } else {
        for (int c = lane; c < H; c += 32)
#pragma unroll
            for (int i = 0; i < kRW; ++i) srow[i * H + c] = 1.f;
    }
    __syncwarp();
    for (int j = lane; j < H3; j += 32) {
        float ai[kRW], ah[kRW];
#pragma unroll
        for (int i = 0; i < kRW; ++i) ai[i] = ah[i] = 0.f;
        for (int k = 0; k < H; ++k) {
            const float wi = Wih_t[k * (H3 + 1) + j];
            const float wh = Whh_t[k * (H3 + 1) + j];
#pragma unroll
            for (int i = 0; i < kRW; ++i) {
                ai[i] = fmaf(wi, xrow[i * H + k], ai[i]);
                ah[i] = fmaf(wh, hrow[i * H + k], ah[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < kRW; ++i) {
            gi[i * H3 + j] = ai[i];
            gh[i * H3 + j] = ah[i];
        }
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < kRW; ++i) {
        if (flags & SPG_GRU_LAYERNORM) {
            float si = 0.f, sh = 0.f;
            for (int j = lane; j < H3; j += 32) {
                si += gi[i * H3 + j];
                sh += gh[i * H3 + j];
            }
            si = warp_sum(si) / (float)H3;
            sh = warp_sum(sh) / (float)H3;
            float vi = 0.f, vh = 0.f;
            for (int j = lane; j < H3; j += 32) {
                const float di = gi[i * H3 + j] - si, dh = gh[i * H3 + j] - sh;
                vi = fmaf(di, di, vi);
                vh = fmaf(dh, dh, vh);
            }
            vi = warp_sum(vi) / (float)H3;
            vh = warp_sum(vh) / (float)H3;
            stats[i][0] = si;
            stats[i][1] = rsqrtf(vi + 1e-5f);
            stats[i][2] = sh;
            stats[i][3] = rsqrtf(vh + 1e-5f);
        } else {
            stats[i][0] = 0.f;
            stats[i][1] = 1.f;
            stats[i][2] = 0.f;
            stats[i][3] = 1.f;
        }
    }
}

template <int kRW>
__global__ void __launch_bounds__(kGruWarps * 32)
gru_fwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
               const float* __restrict__ w_ih, const float* __restrict__ w_hh,
               const float* __restrict__ b_ih, const float* __restrict__ b_hh,
               const float* __restrict__ w_ig, const float* __restrict__ b_ig,
               float* __restrict__ hy, int64_t n_rows, int H, int flags) {
    extern __shared__ float sm[];
    gru_load_weights(sm, w_ih, w_hh, w_ig, H, flags & SPG_GRU_INGATE);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* scratch = sm + gru_weight_floats(H) + warp * gru_scratch_floats(H, kRW);
    const float* hrow = scratch;
    const float* gi = scratch + 3 * kRW * H;
    const float* gh = gi + kRW * 3 * H;
    const int H3 = 3 * H;
    const bool has_bias = flags & SPG_GRU_BIAS;
    const int64_t warps_total = (int64_t)gridDim.x * kGruWarps;
    for (int64_t row0 = ((int64_t)blockIdx.x * kGruWarps + warp) * kRW; row0 < n_rows;
         row0 += warps_total * kRW) {
        float st[kRW][4];
        gru_rows_forward<kRW>(sm, scratch, H, flags, x, h, b_ig, row0, n_rows, lane, st);
        for (int c = lane; c < H; c += 32) {
            const float bir = has_bias ? b_ih[c] : 0.f, biz = has_bias ? b_ih[H + c] : 0.f,
                        bin = has_bias ? b_ih[2 * H + c] : 0.f;
            const float bhr = has_bias ? b_hh[c] : 0.f, bhz = has_bias ? b_hh[H + c] : 0.f,
                        bhn = has_bias ? b_hh[2 * H + c] : 0.f;
#pragma unroll
            for (int i = 0; i < kRW; ++i) {
                const int64_t row = row0 + i;
                if (row >= n_rows) break;
                const float i_r = (gi[i * H3 + c] - st[i][0]) * st[i][1];
                const float i_z = (gi[i * H3 + H + c] - st[i][0]) * st[i][1];
                const float i_n = (gi[i * H3 + 2 * H + c] - st[i][0]) * st[i][1];
                const float h_r = (gh[i * H3 + c] - st[i][2]) * st[i][3];
                const float h_z = (gh[i * H3 + H + c] - st[i][2]) * st[i][3];
                const float h_n = (gh[i * H3 + 2 * H + c] - st[i][2]) * st[i][3];
                const float rg = sigmoidf_(i_r + bir + h_r + bhr);
                const float zg = sigmoidf_(i_z + biz + h_z + bhz);
                const float ng = tanhf(i_n + bin + rg * (h_n + bhn));
                const float hv = hrow[i * H + c];
                hy[row * H + c] = ng + zg * (hv - ng);
            }
        }
        __syncwarp();
    }
}

template <int NU, int kRW>
__global__ void __launch_bounds__(kGruWarps * 32)
gru_bwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
               const float* __restrict__ gy, const float* __restrict__ w_ih,
               const float* __restrict__ w_hh, const float* __restrict__ b_ih,
               const float* __restrict__ b_hh, const float* __restrict__ w_ig,
               const float* __restrict__ b_ig, float* __restrict__ d_x, float* __restrict__ d_h,
               float* __restrict__ d_gi_out, float* __restrict__ d_gh_out,
               float* __restrict__ d_q_out, float* __restrict__ xprime_out,
               float* __restrict__ dpre_out, int64_t n_rows, int H, int flags) {
    extern __shared__ float sm[];
    gru_load_weights(sm, w_ih, w_hh, w_ig, H, flags & SPG_GRU_INGATE);
    __syncthreads();
    const float* Wig_t = sm;
    const float* Wih_t = Wig_t + H * (H + 1);
    const float* Whh_t = Wih_t + H * (3 * H + 1);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* scratch = sm + gru_weight_floats(H) + warp * gru_scratch_floats(H, kRW);
    float* hrow = scratch;
    float* xrow = hrow + kRW * H;   // x' (gated input)
    float* srow = xrow + kRW * H;   // sigmoid(q); reused below for d_q
    float* gi = srow + kRW * H;     // raw gate inputs -> overwritten with d_gi
    float* gh = gi + kRW * 3 * H;   // raw gate inputs -> overwritten with d_gh
    const int H3 = 3 * H;
    const bool has_bias = flags & SPG_GRU_BIAS;
    const bool ln = flags & SPG_GRU_LAYERNORM;
    const bool ingate = flags & SPG_GRU_INGATE;
    const int64_t warps_total = (int64_t)gridDim.x * kGruWarps;
    for (int64_t row0 = ((int64_t)blockIdx.x * kGruWarps + warp) * kRW; row0 < n_rows;
         row0 += warps_total * kRW) {
        float st[kRW][4];
        gru_rows_forward<kRW>(sm, scratch, H, flags, x, h, b_ig, row0, n_rows, lane, st);
        // ---- gate gradients (w.r.t. the normalised gate inputs), in place over gi/gh
        float dh_direct[kRW][NU];  // column c = lane + 32*u
#pragma unroll
        for (int i = 0; i < kRW; ++i) {
            const int64_t row = row0 + i;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int c = lane + 32 * u;
                if (c >= H) continue;
                const float bir = has_bias ? b_ih[c] : 0.f, biz = has_bias ? b_ih[H + c] : 0.f,
                            bin = has_bias ? b_ih[2 * H + c] : 0.f;
                const float bhr = has_bias ? b_hh[c] : 0.f, bhz = has_bias ? b_hh[H + c] : 0.f,
                            bhn = has_bias ? b_hh[2 * H + c] : 0.f;
                const float i_r = (gi[i * H3 + c] - st[i][0]) * st[i][1];
                const float i_z = (gi[i * H3 + H + c] - st[i][0]) * st[i][1];
                const float i_n = (gi[i * H3 + 2 * H + c] - st[i][0]) * st[i][1];
                const float h_r = (gh[i * H3 + c] - st[i][2]) * st[i][3];
                const float h_z = (gh[i * H3 + H + c] - st[i][2]) * st[i][3];
                const float h_n = (gh[i * H3 + 2 * H + c] - st[i][2]) * st[i][3];
                const float rg = sigmoidf_(i_r + bir + h_r + bhr);
                const float zg = sigmoidf_(i_z + biz + h_z + bhz);
                const float ng = tanhf(i_n + bin + rg * (h_n + bhn));
                const float hv = hrow[i * H + c];
                const float g = row < n_rows ? gy[row * H + c] : 0.f;
                const float d_n = g * (1.f - zg);
                const float d_z = g * (hv - ng);
                dh_direct[i][u] = g * zg;
                const float d_pn = d_n * (1.f - ng * ng);
                const float d_r = d_pn * (h_n + bhn);
                const float d_pz = d_z * zg * (1.f - zg);
                const float d_pr = d_r * rg * (1.f - rg);
                if (row < n_rows) {
                    float* dp = dpre_out + row * 4 * H;
                    dp[c] = d_pr;
                    dp[H + c] = d_pz;
                    dp[2 * H + c] = d_pn;
                    dp[3 * H + c] = d_pn * rg;
                }
                // y-hat (normalised value) is needed by the norm backward: keep it in
                // registers via recomputation below; store dy now, y-hat products later.
                // Layout trick: write dy into gi/gh only after the lane has read all of
                // its own entries (each lane owns columns c, H+c, 2H+c of both arrays).
                gi[i * H3 + c] = ln ? i_r : 0.f;          // stash y-hat
                gi[i * H3 + H + c] = ln ? i_z : 0.f;
                gi[i * H3 + 2 * H + c] = ln ? i_n : 0.f;
                gh[i * H3 + c] = ln ? h_r : 0.f;
                gh[i * H3 + H + c] = ln ? h_z : 0.f;
                gh[i * H3 + 2 * H + c] = ln ? h_n : 0.f;
                // dy kept in registers through the second scratch: reuse srow? no - it
                // still holds sigmoid(q).  Use dpre_out (global, just written) instead.
            }
        }
        __syncwarp();
        // ---- layer-norm backward: d_u = rstd * (dy - mean(dy) - yhat*mean(dy*yhat))
#pragma unroll
        for (int i = 0; i < kRW; ++i) {
            const int64_t row = row0 + i;
            const bool live = row < n_rows;
            const float* dp = dpre_out + (live ? row : 0) * 4 * H;
            float m1i = 0.f, m2i = 0.f, m1h = 0.f, m2h = 0.f;
            if (ln) {
                for (int j = lane; j < H3; j += 32) {
                    const float dyi = live ? dp[j] : 0.f;
                    const float dyh = live ? (j < 2 * H ? dp[j] : dp[j + H]) : 0.f;
                    m1i += dyi;
                    m2i = fmaf(dyi, gi[i * H3 + j], m2i);
                    m1h += dyh;
                    m2h = fmaf(dyh, gh[i * H3 + j], m2h);
                }
                m1i = warp_sum(m1i) / (float)H3;
                m2i = warp_sum(m2i) / (float)H3;
                m1h = warp_sum(m1h) / (float)H3;
                m2h = warp_sum(m2h) / (float)H3;
            }
            for (int j = lane; j < H3; j += 32) {
                const float dyi = live ? dp[j] : 0.f;
                const float dyh = live ? (j < 2 * H ? dp[j] : dp[j + H]) : 0.f;
                float dui, duh;
                if (ln) {
                    dui = st[i][1] * (dyi - m1i - gi[i * H3 + j] * m2i);
                    duh = st[i][3] * (dyh - m1h - gh[i * H3 + j] * m2h);
                } else {
                    dui = dyi;
                    duh = dyh;
                }
                gi[i * H3 + j] = dui;
                gh[i * H3 + j] = duh;
                if (live) {
                    d_gi_out[row * H3 + j] = dui;
                    d_gh_out[row * H3 + j] = duh;
                }
            }
        }
        __syncwarp();
        // ---- d_x' = d_gi * W_ih ; d_h += d_gh * W_hh
        {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int k = lane + 32 * u;
                if (k >= H) continue;
                float ax[kRW], ah[kRW];
#pragma unroll
                for (int i = 0; i < kRW; ++i) ax[i] = ah[i] = 0.f;
                for (int j = 0; j < H3; ++j) {
                    const float wi = Wih_t[k * (H3 + 1) + j];
                    const float wh = Whh_t[k * (H3 + 1) + j];
#pragma unroll
                    for (int i = 0; i < kRW; ++i) {
                        ax[i] = fmaf(wi, gi[i * H3 + j], ax[i]);
                        ah[i] = fmaf(wh, gh[i * H3 + j], ah[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < kRW; ++i) {
                    const int64_t row = row0 + i;
                    const float sg = srow[i * H + k];
                    const float xp = xrow[i * H + k];       // x' = s*x
                    float dq = 0.f;
                    float dxv = ax[i];
                    if (ingate) {
                        // x = x'/s is not safe when s underflows: reload the raw input.
                        const float xin = row < n_rows ? x[row * H + k] : 0.f;
                        const float ds = ax[i] * xin;
                        dxv = ax[i] * sg;
                        dq = ds * sg * (1.f - sg);
                    }
                    dh_direct[i][u] += ah[i];
                    if (row < n_rows) {
                        d_x[row * H + k] = dxv;
                        xprime_out[row * H + k] = xp;
                        d_q_out[row * H + k] = dq;
                    }
                    srow[i * H + k] = dq;  // only this lane touches srow[.][k]
                }
            }
        }
        __syncwarp();
        // ---- d_h += d_q * W_ig
        {
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const int k = lane + 32 * u;
                if (k >= H) continue;
                float a[kRW];
#pragma unroll
                for (int i = 0; i < kRW; ++i) a[i] = 0.f;
                if (ingate) {
                    for (int c = 0; c < H; ++c) {
                        const float wv = Wig_t[k * (H + 1) + c];
#pragma unroll
                        for (int i = 0; i < kRW; ++i) a[i] = fmaf(wv, srow[i * H + c], a[i]);
                    }
                }
#pragma unroll
                for (int i = 0; i < kRW; ++i) {
                    const int64_t row = row0 + i;
                    if (row < n_rows) d_h[row * H + k] = dh_direct[i][u] + a[i];
                }
            }
        }
        __syncwarp();
    }
}

static inline size_t gru_smem_bytes(int H, int rw) {
    return sizeof(float) * ((size_t)gru_weight_floats(H) + (size_t)kGruWarps * gru_scratch_floats(H, rw));
}

// rows per warp: 4 amortises the shared-memory weight reads when there are enough rows to fill
// the GPU; small graphs (the S3DIS training batches) use 1 so that every row gets its own warp.
static inline int gru_rows_per_warp(int64_t n_rows) {
    return n_rows >= (int64_t)kNumSMs * kGruWarps * 4 * 2 ? 4 : 1;
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_gru_fwd(const float* x, const float* h, const float* weight_ih, const float* weight_hh,
                const float* bias_ih, const float* bias_hh, const float* ig_weight,
                const float* ig_bias, float* hy, int64_t n_rows, int hidden, int flags,
                spg_stream_t stream) {
    if (n_rows < 0 || hidden <= 0) return SPG_E_BADARG;
    if (n_rows == 0) return SPG_OK;
    if (!x || !h || !weight_ih || !weight_hh || !hy) return SPG_E_BADARG;
    if ((flags & SPG_GRU_BIAS) && (!bias_ih || !bias_hh)) return SPG_E_BADARG;
    if ((flags & SPG_GRU_INGATE) && (!ig_weight || !ig_bias)) return SPG_E_BADARG;
    const int rw = gru_rows_per_warp(n_rows);
    const size_t smem = gru_smem_bytes(hidden, rw);
    if (hidden > 128 || smem > 227 * 1024) return SPG_E_UNSUPPORTED;
    int64_t blocks = ceil_div64(n_rows, (int64_t)kGruWarps * rw);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
#define SPG_GRU_FWD_CASE(RW)                                                                      \
    {                                                                                             \
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_kernel<RW>,                                  \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                             (int)smem);                                          \
        if (e != cudaSuccess) return (int)e;                                                      \
        SPG_LAUNCH(K_GRU_FWD, (cudaStream_t)stream, gru_fwd_kernel<RW>, (unsigned)blocks,         \
                   kGruWarps * 32, smem, x, h, weight_ih, weight_hh, bias_ih, bias_hh, ig_weight, \
                   ig_bias, hy, n_rows, hidden, flags);                                           \
    }
    if (rw == 4) { SPG_GRU_FWD_CASE(4) } else { SPG_GRU_FWD_CASE(1) }
#undef SPG_GRU_FWD_CASE
    return launch_status();
}

int spg_gru_bwd(const float* x, const float* h, const float* grad_hy, const float* weight_ih,
                const float* weight_hh, const float* bias_ih, const float* bias_hh,
                const float* ig_weight, const float* ig_bias, float* d_x, float* d_h,
                float* d_gi, float* d_gh, float* d_q, float* xprime, float* dpre,
                int64_t n_rows, int hidden, int flags, spg_stream_t stream) {
    if (n_rows < 0 || hidden <= 0) return SPG_E_BADARG;
    if (n_rows == 0) return SPG_OK;
    if (!x || !h || !grad_hy || !weight_ih || !weight_hh || !d_x || !d_h || !d_gi || !d_gh ||
        !d_q || !xprime || !dpre)
        return SPG_E_BADARG;
    if ((flags & SPG_GRU_BIAS) && (!bias_ih || !bias_hh)) return SPG_E_BADARG;
    if ((flags & SPG_GRU_INGATE) && (!ig_weight || !ig_bias)) return SPG_E_BADARG;
    const int rw = gru_rows_per_warp(n_rows);
    const size_t smem = gru_smem_bytes(hidden, rw);
    if (hidden > 128 || smem > 227 * 1024) return SPG_E_UNSUPPORTED;
    int64_t blocks = ceil_div64(n_rows, (int64_t)kGruWarps * rw);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
#define SPG_GRU_BWD_CASE(NU) { SPG_GRU_BWD_CASE2(NU, 4) else SPG_GRU_BWD_CASE2(NU, 1) }
#define SPG_GRU_BWD_CASE2(NU, RW) if (rw == RW)                                                                      \
    {                                                                                             \
        cudaError_t e = cudaFuncSetAttribute(gru_bwd_kernel<NU, RW>,                                  \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,         \
                                             (int)smem);                                          \
        if (e != cudaSuccess) return (int)e;                                                      \
        SPG_LAUNCH(K_GRU_BWD, (cudaStream_t)stream, (gru_bwd_kernel<NU, RW>), (unsigned)blocks,         \
                   kGruWarps * 32, smem, x, h, grad_hy, weight_ih, weight_hh, bias_ih, bias_hh,   \
                   ig_weight, ig_bias, d_x, d_h, d_gi, d_gh, d_q, xprime, dpre, n_rows, hidden,   \
                   flags);                                                                        \
    }
    if (hidden <= 32) SPG_GRU_BWD_CASE(1)
    else if (hidden <= 64) SPG_GRU_BWD_CASE(2)
    else SPG_GRU_BWD_CASE(4)
#undef SPG_GRU_BWD_CASE
#undef SPG_GRU_BWD_CASE2
    return launch_status();
}

}  // extern "C"
