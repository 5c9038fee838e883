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
                                                 const float* x, const float* h,
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

// New hidden state of the RW rows whose gate inputs gru_rows_forward left in `scratch`.
template <int kRW>
__device__ __forceinline__ void gru_rows_emit(const float* scratch, int H, int flags,
                                              const float* __restrict__ b_ih,
                                              const float* __restrict__ b_hh, float* hy,
                                              int64_t row0, int64_t n_rows, int lane,
                                              const float st[kRW][4]) {
    const float* hrow = scratch;
    const float* gi = scratch + 3 * kRW * H;
    const float* gh = gi + kRW * 3 * H;
    const int H3 = 3 * H;
    const bool has_bias = flags & SPG_GRU_BIAS;
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

template <int kRW>
__global__ void __launch_bounds__(kGruWarps * 32)
gru_fwd_kernel(const float* __restrict__ x, const float* __restrict__ h,
               const float* __restrict__ w_ih, const float* __restrict__ w_hh,
               const float* __restrict__ b_ih, const float* __restrict__ b_hh,
               const float* __restrict__ w_ig, const float* __restrict__ b_ig,
               float* __restrict__ hy, int64_t n_rows, int H, int flags) {
    SPG_PDL_ENTRY();
    extern __shared__ float sm[];
    gru_load_weights(sm, w_ih, w_hh, w_ig, H, flags & SPG_GRU_INGATE);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* scratch = sm + gru_weight_floats(H) + warp * gru_scratch_floats(H, kRW);
    const int64_t warps_total = (int64_t)gridDim.x * kGruWarps;
    for (int64_t row0 = ((int64_t)blockIdx.x * kGruWarps + warp) * kRW; row0 < n_rows;
         row0 += warps_total * kRW) {
        float st[kRW][4];
        gru_rows_forward<kRW>(sm, scratch, H, flags, x, h, b_ig, row0, n_rows, lane, st);
        gru_rows_emit<kRW>(scratch, H, flags, b_ih, b_hh, hy, row0, n_rows, lane, st);
    }
}

// Backward of the cell for the RW rows starting at row0 (one warp).  x, h and gy may have been
// written earlier in the same kernel (fused recurrent kernels), hence no __restrict__ on them.
template <int NU, int kRW>
__device__ __forceinline__ void gru_rows_backward(
    const float* sm, float* scratch, int H, int flags, const float* x, const float* h,
    const float* gy, const float* __restrict__ b_ih, const float* __restrict__ b_hh,
    const float* __restrict__ b_ig, float* d_x, float* d_h, float* __restrict__ d_gi_out,
    float* __restrict__ d_gh_out, float* __restrict__ d_q_out, float* __restrict__ xprime_out,
    float* dpre_out, int64_t row0, int64_t n_rows, int lane) {
    const float* Wig_t = sm;
    const float* Wih_t = Wig_t + H * (H + 1);
    const float* Whh_t = Wih_t + H * (3 * H + 1);
    float* hrow = scratch;
    float* xrow = hrow + kRW * H;   // x' (gated input)
    float* srow = xrow + kRW * H;   // sigmoid(q); reused below for d_q
    float* gi = srow + kRW * H;     // raw gate inputs -> overwritten with d_gi
    float* gh = gi + kRW * 3 * H;   // raw gate inputs -> overwritten with d_gh
    const int H3 = 3 * H;
    const bool has_bias = flags & SPG_GRU_BIAS;
    const bool ln = flags & SPG_GRU_LAYERNORM;
    const bool ingate = flags & SPG_GRU_INGATE;
    {
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
    SPG_PDL_ENTRY();
    extern __shared__ float sm[];
    gru_load_weights(sm, w_ih, w_hh, w_ig, H, flags & SPG_GRU_INGATE);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* scratch = sm + gru_weight_floats(H) + warp * gru_scratch_floats(H, kRW);
    const int64_t warps_total = (int64_t)gridDim.x * kGruWarps;
    for (int64_t row0 = ((int64_t)blockIdx.x * kGruWarps + warp) * kRW; row0 < n_rows;
         row0 += warps_total * kRW)
        gru_rows_backward<NU, kRW>(sm, scratch, H, flags, x, h, gy, b_ih, b_hh, b_ig, d_x, d_h,
                                   d_gi_out, d_gh_out, d_q_out, xprime_out, dpre_out, row0, n_rows,
                                   lane);
}


// ------------------------------------------------------------------ fused recurrence
// The R x {ECC, cell} loop of RNNGraphConvModule (ref: learning/modules.py:160-180) as ONE
// persistent kernel each way, for the training-batch regime (a few thousand superpoints) where
// 2R..3R separate launches are pure latency.  A warp owns a node for the whole recurrence:
//   forward   r:  inp_i = ECC(h_r)_i  ->  h_{r+1,i} = cell(inp_i, h_{r,i})   | grid barrier
//   backward  r:  (d_inp_i, d_h_i) = cell'(g_i)  | grid barrier |  g_i = d_h_i + ECC'(d_inp)_i (+cat)
// so only the neighbour exchange crosses the barrier; the cell weights are loaded into shared
// memory once per CTA instead of once per step.  Vector filters, H = 32, fp32, no idxe.
constexpr int kRecH = 32;

// All CTAs are co-resident (the launchers cap the grid with the occupancy API); `counter` counts
// arrivals monotonically and is zeroed by the launcher.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned v;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
        } while (v < target);
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ float4 ldcg4(const float* p) {
    return __ldcg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float4 slot_reduce(float4 acc) {
#pragma unroll
    for (int o = 8; o <= 16; o <<= 1) {
        acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
        acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o);
        acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
    }
    return acc;
}

__global__ void __launch_bounds__(kGruWarps * 32)
rnn_vv_fwd_kernel(float* hs, float* inps, const float4* __restrict__ w,
                  const int* __restrict__ rowptr, const int* __restrict__ idxn,
                  const float* __restrict__ w_ih, const float* __restrict__ w_hh,
                  const float* __restrict__ b_ih, const float* __restrict__ b_hh,
                  const float* __restrict__ w_ig, const float* __restrict__ b_ig, int n, int R,
                  int flags, unsigned* barrier) {
    SPG_PDL_ENTRY();
    extern __shared__ float sm[];
    gru_load_weights(sm, w_ih, w_hh, w_ig, kRecH, flags & SPG_GRU_INGATE);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = lane >> 3, sub = lane & 7;
    float* scratch = sm + gru_weight_floats(kRecH) + warp * gru_scratch_floats(kRecH, 1);
    const int gwarp = blockIdx.x * kGruWarps + warp, nwarps = gridDim.x * kGruWarps;
    for (int r = 0; r < R; ++r) {
        float* hcur = hs + (size_t)r * n * kRecH;
        float* hnext = hcur + (size_t)n * kRecH;
        float* inp = inps + (size_t)r * n * kRecH;
        for (int node = gwarp; node < n; node += nwarps) {
            const int beg = rowptr[node], end = rowptr[node + 1];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            int e = beg + slot;
            for (; e + 4 < end; e += 8) {
                const int s0 = __ldg(idxn + e), s1 = __ldg(idxn + e + 4);
                const float4 w0 = __ldg(w + (int64_t)e * 8 + sub);
                const float4 w1 = __ldg(w + (int64_t)(e + 4) * 8 + sub);
                const float4 x0 = ldcg4(hcur + (int64_t)s0 * kRecH + sub * 4);
                const float4 x1 = ldcg4(hcur + (int64_t)s1 * kRecH + sub * 4);
                acc.x = fmaf(x0.x, w0.x, acc.x); acc.y = fmaf(x0.y, w0.y, acc.y);
                acc.z = fmaf(x0.z, w0.z, acc.z); acc.w = fmaf(x0.w, w0.w, acc.w);
                acc.x = fmaf(x1.x, w1.x, acc.x); acc.y = fmaf(x1.y, w1.y, acc.y);
                acc.z = fmaf(x1.z, w1.z, acc.z); acc.w = fmaf(x1.w, w1.w, acc.w);
            }
            if (e < end) {
                const int s0 = __ldg(idxn + e);
                const float4 w0 = __ldg(w + (int64_t)e * 8 + sub);
                const float4 x0 = ldcg4(hcur + (int64_t)s0 * kRecH + sub * 4);
                acc.x = fmaf(x0.x, w0.x, acc.x); acc.y = fmaf(x0.y, w0.y, acc.y);
                acc.z = fmaf(x0.z, w0.z, acc.z); acc.w = fmaf(x0.w, w0.w, acc.w);
            }
            acc = slot_reduce(acc);
            if (slot == 0) {
                const int deg = end - beg;
                if (deg > 0) {
                    const float d = (float)deg;
                    acc.x /= d; acc.y /= d; acc.z /= d; acc.w /= d;
                }
                *reinterpret_cast<float4*>(inp + (int64_t)node * kRecH + sub * 4) = acc;
            }
            __syncwarp();
            float st[1][4];
            gru_rows_forward<1>(sm, scratch, kRecH, flags, inp, hcur, b_ig, node, n, lane, st);
            gru_rows_emit<1>(scratch, kRecH, flags, b_ih, b_hh, hnext, node, n, lane, st);
        }
        if (r + 1 < R) grid_barrier(barrier, (unsigned)(r + 1) * gridDim.x);
    }
}

__global__ void __launch_bounds__(kGruWarps * 32)
rnn_vv_bwd_kernel(const float* __restrict__ hs, const float* __restrict__ inps,
                  const float4* __restrict__ w, const float* __restrict__ gtop,
                  const float* __restrict__ gcat, const int* __restrict__ tgt_rowptr,
                  const int* __restrict__ src_rowptr, const int* __restrict__ src_perm,
                  const int* __restrict__ edge_tgt, const float* __restrict__ w_ih,
                  const float* __restrict__ w_hh, const float* __restrict__ b_ih,
                  const float* __restrict__ b_hh, const float* __restrict__ w_ig,
                  const float* __restrict__ b_ig, float* ginp, float* dh, float* gh,
                  float* __restrict__ d_gi, float* __restrict__ d_gh, float* __restrict__ d_q,
                  float* __restrict__ xp, float* dpre, int n, int R, int flags,
                  unsigned* barrier) {
    SPG_PDL_ENTRY();
    extern __shared__ float sm[];
    gru_load_weights(sm, w_ih, w_hh, w_ig, kRecH, flags & SPG_GRU_INGATE);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int slot = lane >> 3, sub = lane & 7;
    float* scratch = sm + gru_weight_floats(kRecH) + warp * gru_scratch_floats(kRecH, 1);
    const int gwarp = blockIdx.x * kGruWarps + warp, nwarps = gridDim.x * kGruWarps;
    const size_t plane = (size_t)n * kRecH;
    for (int r = R - 1; r >= 0; --r) {
        const float* gy = (r == R - 1) ? gtop : gh;   // gh rows are produced by the same warp
        float* ginp_r = ginp + r * plane;
        for (int node = gwarp; node < n; node += nwarps)
            gru_rows_backward<1, 1>(sm, scratch, kRecH, flags, inps + r * plane, hs + r * plane, gy,
                                    b_ih, b_hh, b_ig, ginp_r, dh, d_gi + 3 * r * plane,
                                    d_gh + 3 * r * plane, d_q + r * plane, xp + r * plane,
                                    dpre + 4 * r * plane, node, n, lane);
        grid_barrier(barrier, (unsigned)(R - r) * gridDim.x);
        for (int node = gwarp; node < n; node += nwarps) {
            const int beg = src_rowptr[node], end = src_rowptr[node + 1];
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = beg + slot; p < end; p += 4) {
                const int e = __ldg(src_perm + p);
                const int tg = __ldg(edge_tgt + e);
                const float4 wv = __ldg(w + (int64_t)e * 8 + sub);
                const float inv = 1.f / (float)(__ldg(tgt_rowptr + tg + 1) - __ldg(tgt_rowptr + tg));
                const float4 gv = ldcg4(ginp_r + (int64_t)tg * kRecH + sub * 4);
                acc.x = fmaf(wv.x, gv.x * inv, acc.x); acc.y = fmaf(wv.y, gv.y * inv, acc.y);
                acc.z = fmaf(wv.z, gv.z * inv, acc.z); acc.w = fmaf(wv.w, gv.w * inv, acc.w);
            }
            acc = slot_reduce(acc);
            if (slot == 0) {
                const float4 a = *reinterpret_cast<const float4*>(dh + (int64_t)node * kRecH + sub * 4);
                acc.x += a.x; acc.y += a.y; acc.z += a.z; acc.w += a.w;
                if (gcat) {
                    const float4 c = __ldg(reinterpret_cast<const float4*>(
                        gcat + r * plane + (int64_t)node * kRecH + sub * 4));
                    acc.x += c.x; acc.y += c.y; acc.z += c.z; acc.w += c.w;
                }
                *reinterpret_cast<float4*>(gh + (int64_t)node * kRecH + sub * 4) = acc;
            }
            __syncwarp();
        }
    }
}

static inline int rnn_grid(const void* kernel, size_t smem, int n) {
    int per_sm = 0, sms = 0, dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kGruWarps * 32, smem) !=
        cudaSuccess)
        return -1;
    if (per_sm > 2) per_sm = 2;
    int64_t blocks = ceil_div64(n, kGruWarps);
    const int64_t cap = (int64_t)per_sm * sms;
    if (blocks > cap) blocks = cap;
    return (int)blocks;   // 0 if the kernel does not fit at all
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

int spg_rnn_vv_supported(int64_t n_nodes, int hidden) {
    // any number of nodes: the kernels walk the nodes grid-strided inside every step
    return hidden == kRecH && n_nodes > 0 && n_nodes < ((int64_t)1 << 31) / (4 * kRecH);
}

int spg_rnn_vv_fwd(float* hs, float* inps, const float* w, const int32_t* tgt_rowptr,
                   const int32_t* idxn, const float* weight_ih, const float* weight_hh,
                   const float* bias_ih, const float* bias_hh, const float* ig_weight,
                   const float* ig_bias, int64_t n_nodes, int hidden, int n_repeats, int flags,
                   void* barrier_ws, spg_stream_t stream) {
    if (n_nodes < 0 || n_repeats < 0) return SPG_E_BADARG;
    if (n_nodes == 0 || n_repeats == 0) return SPG_OK;
    if (!spg_rnn_vv_supported(n_nodes, hidden)) return SPG_E_UNSUPPORTED;
    if (!hs || !inps || !w || !tgt_rowptr || !idxn || !weight_ih || !weight_hh || !barrier_ws)
        return SPG_E_BADARG;
    if ((flags & SPG_GRU_BIAS) && (!bias_ih || !bias_hh)) return SPG_E_BADARG;
    if ((flags & SPG_GRU_INGATE) && (!ig_weight || !ig_bias)) return SPG_E_BADARG;
    const size_t smem = gru_smem_bytes(hidden, 1);
    cudaError_t e = cudaFuncSetAttribute(rnn_vv_fwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int blocks = rnn_grid((const void*)rnn_vv_fwd_kernel, smem, (int)n_nodes);
    if (blocks <= 0) return SPG_E_UNSUPPORTED;
    e = cudaMemsetAsync(barrier_ws, 0, sizeof(unsigned), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    // cooperative launch: the driver guarantees that all CTAs are co-resident (the grid barrier
    // spins on them) or fails the launch; other streams' kernels cannot starve the grid
    const float4* w4 = reinterpret_cast<const float4*>(w);
    int n_i = (int)n_nodes;
    unsigned* bar = reinterpret_cast<unsigned*>(barrier_ws);
    void* kargs[] = {&hs, &inps, &w4, &tgt_rowptr, &idxn, &weight_ih, &weight_hh, &bias_ih, &bias_hh,
                     &ig_weight, &ig_bias, &n_i, &n_repeats, &flags, &bar};
    {
        ::spg::LaunchScope _scope(K_RNN_FWD, (cudaStream_t)stream);
        e = cudaLaunchCooperativeKernel((const void*)rnn_vv_fwd_kernel, dim3((unsigned)blocks),
                                        dim3(kGruWarps * 32), kargs, smem, (cudaStream_t)stream);
    }
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}

int spg_rnn_vv_bwd(const float* hs, const float* inps, const float* w, const float* grad_top,
                   const float* grad_cat, const int32_t* tgt_rowptr, const int32_t* src_rowptr,
                   const int32_t* src_perm, const int32_t* edge_tgt, const float* weight_ih,
                   const float* weight_hh, const float* bias_ih, const float* bias_hh,
                   const float* ig_weight, const float* ig_bias, float* grad_inp, float* d_h_ws,
                   float* grad_h0, float* d_gi, float* d_gh, float* d_q, float* xprime,
                   float* dpre, int64_t n_nodes, int hidden, int n_repeats, int flags,
                   void* barrier_ws, spg_stream_t stream) {
    if (n_nodes < 0 || n_repeats < 0) return SPG_E_BADARG;
    if (n_nodes == 0 || n_repeats == 0) return SPG_OK;
    if (!spg_rnn_vv_supported(n_nodes, hidden)) return SPG_E_UNSUPPORTED;
    if (!hs || !inps || !w || !grad_top || !tgt_rowptr || !src_rowptr || !src_perm || !edge_tgt ||
        !weight_ih || !weight_hh || !grad_inp || !d_h_ws || !grad_h0 || !d_gi || !d_gh || !d_q ||
        !xprime || !dpre || !barrier_ws)
        return SPG_E_BADARG;
    if ((flags & SPG_GRU_BIAS) && (!bias_ih || !bias_hh)) return SPG_E_BADARG;
    if ((flags & SPG_GRU_INGATE) && (!ig_weight || !ig_bias)) return SPG_E_BADARG;
    const size_t smem = gru_smem_bytes(hidden, 1);
    cudaError_t e = cudaFuncSetAttribute(rnn_vv_bwd_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    const int blocks = rnn_grid((const void*)rnn_vv_bwd_kernel, smem, (int)n_nodes);
    if (blocks <= 0) return SPG_E_UNSUPPORTED;
    e = cudaMemsetAsync(barrier_ws, 0, sizeof(unsigned), (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    const float4* w4 = reinterpret_cast<const float4*>(w);
    int n_i = (int)n_nodes;
    unsigned* bar = reinterpret_cast<unsigned*>(barrier_ws);
    void* kargs[] = {&hs, &inps, &w4, &grad_top, &grad_cat, &tgt_rowptr, &src_rowptr, &src_perm,
                     &edge_tgt, &weight_ih, &weight_hh, &bias_ih, &bias_hh, &ig_weight, &ig_bias,
                     &grad_inp, &d_h_ws, &grad_h0, &d_gi, &d_gh, &d_q, &xprime, &dpre, &n_i,
                     &n_repeats, &flags, &bar};
    {
        ::spg::LaunchScope _scope(K_RNN_BWD, (cudaStream_t)stream);
        e = cudaLaunchCooperativeKernel((const void*)rnn_vv_bwd_kernel, dim3((unsigned)blocks),
                                        dim3(kGruWarps * 32), kargs, smem, (cudaStream_t)stream);
    }
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}

}  // extern "C"
