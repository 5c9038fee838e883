// Step glue on the hot path: weighted cross entropy (+ its gradient) and the fused
// gradient clamp + Adam update over one flat parameter buffer.
//
// Reference semantics: learning/main.py:205 (nn.functional.cross_entropy with class
// weights, ignore_index -100), :210-213 (element-wise clamp of every gradient, then
// optimizer.step()).  The reference launches several kernels per parameter tensor
// (~60 tensors); here the whole model is one flat buffer and one launch.
#include "common.cuh"

namespace spg {

// One warp per row.  acc[0] += w*nll, acc[1] += w (double atomics).
// d_logits = w * (softmax - onehot), normalised by the second kernel.
__global__ void __launch_bounds__(256)
ce_loss_kernel(const float* __restrict__ logits, const int64_t* __restrict__ target,
               const float* __restrict__ cw, int64_t ignore_index, double* __restrict__ acc,
               float* __restrict__ dlogits, int64_t n, int C) {
    SPG_PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= n) return;
    const int64_t y = target[row];
    const float* src = logits + row * C;
    float* dst = dlogits ? dlogits + row * C : nullptr;
    if (y == ignore_index || y < 0 || y >= C) {
        if (dst)
            for (int c = lane; c < C; c += 32) dst[c] = 0.f;
        return;
    }
    float mx = -3.4e38f;
    for (int c = lane; c < C; c += 32) mx = fmaxf(mx, src[c]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float se = 0.f;
    for (int c = lane; c < C; c += 32) se += expf(src[c] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    const float w = cw ? cw[y] : 1.f;
    if (dst)
        for (int c = lane; c < C; c += 32)
            dst[c] = w * (expf(src[c] - lse) - (c == y ? 1.f : 0.f));
    if (lane == 0) {
        atomicAdd(acc, (double)(w * (lse - src[y])));
        atomicAdd(acc + 1, (double)w);
    }
}

__global__ void ce_loss_final_kernel(const double* __restrict__ acc, float* __restrict__ loss,
                                     float* __restrict__ dlogits, int64_t total) {
    SPG_PDL_ENTRY();
    const double wsum = acc[1];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) loss[0] = (float)(acc[0] / wsum);  // 0/0 = NaN for an all-ignored batch, as torch
    // an all-ignored batch (wsum == 0) has a NaN loss but ZERO gradients in torch's cross_entropy:
    // every d_logits row was written as 0 by the first kernel, leave it
    if (dlogits && i < total && wsum != 0.0) dlogits[i] = (float)((double)dlogits[i] / wsum);
}

__global__ void __launch_bounds__(256)
clamp_adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                  float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                  float wd, float clip, float gscale, float bc1, float bc2_sqrt) {
    SPG_PDL_ENTRY();
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        // p.grad.clamp_ propagates NaN (fminf/fmaxf would turn it into -clip and hide a divergence)
        if (clip > 0.f && gi == gi) gi = fminf(fmaxf(gi, -clip), clip);
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

// Same update with the step count read from device memory (CUDA-graph friendly: nothing about the
// step is baked into the launch parameters).  step = *counter + 1.
__global__ void __launch_bounds__(256)
clamp_adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                      float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps,
                      float wd, float clip, float gscale, const long long* __restrict__ counter) {
    SPG_PDL_ENTRY();
    const double step = (double)(counter[0] + 1);
    const float bc1 = (float)(1.0 - pow((double)b1, step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
        float gi = g[i] * gscale;
        // p.grad.clamp_ propagates NaN (fminf/fmaxf would turn it into -clip and hide a divergence)
        if (clip > 0.f && gi == gi) gi = fminf(fmaxf(gi, -clip), clip);
        const float pi = p[i];
        if (wd != 0.f) gi = fmaf(wd, pi, gi);
        const float mi = b1 * m[i] + (1.f - b1) * gi;
        const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + eps;
        p[i] = pi - (lr / bc1) * (mi / denom);
    }
}

__global__ void counter_increment_kernel(long long* counter) {
    SPG_PDL_ENTRY(); counter[0] += 1; }

// ---- the step's only collective, fused with what follows it (main.py:210-213) ----------
// One-shot all-reduce over NVLink peer memory + average + element-wise clamp + Adam in ONE kernel.
// Every rank owns a symmetric STAGING buffer of two halves (peer pointers in `peer_stage`).  Block b
//   0. copies its slice of the local flat gradient into half (epoch & 1) of its own staging buffer,
//   1. tells block b of every peer "my slice is there" (release flag, system scope) and waits for theirs,
//   2. reads the same slice from every peer's half over NVLink, sums in rank order (deterministic and
//      bit-identical on all ranks: the replicas cannot drift), scales by 1/world, clamps, applies Adam.
// There is no second ("done reading") handshake: a rank overwrites half h again two launches later, and it
// cannot finish the launch in between before every peer has STARTED that launch, i.e. finished this one
// (stream order) and with it all reads of half h.  0.85 MB per rank: latency-, not bandwidth-bound, hence one-shot
// and no NCCL launch.  Flags carry a launch epoch kept in device memory, so the kernel can sit inside a CUDA
// graph.  Blocks only ever wait for same-index blocks of PEERS, never for local blocks: no co-residency
// requirement.
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(512)
allreduce_clamp_adam_kernel(const float* __restrict__ grad_local, float* const* __restrict__ peer_stage,
                            unsigned* const* __restrict__ peer_flags, int rank, int world, float* __restrict__ p,
                            float* __restrict__ m, float* __restrict__ v, int64_t n, int64_t half_stride, float lr,
                            float b1, float b2, float eps, float wd, float clip, float gscale,
                            long long* __restrict__ counter, unsigned* __restrict__ epoch_ptr,
                            unsigned* __restrict__ ticket) {
    SPG_PDL_ENTRY();
    const unsigned epoch = epoch_ptr[0] + 1u;
    const unsigned nb = gridDim.x;
    const int64_t half = (int64_t)(epoch & 1u) * half_stride;
    unsigned* my_flags = peer_flags[rank];
    float* my_stage = peer_stage[rank] + half;
    const int64_t n4 = n >> 2;
    // step 0: my slice into my staging half
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)nb * blockDim.x)
        reinterpret_cast<float4*>(my_stage)[i] = reinterpret_cast<const float4*>(grad_local)[i];
    if (blockIdx.x == 0)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) my_stage[i] = grad_local[i];
    __syncthreads();
    // step 1: "my slice is complete" to block b of every peer; wait for the same from every peer
    if (threadIdx.x < world) {
        const int peer = threadIdx.x;
        __threadfence_system();
        st_release_sys(peer_flags[peer] + (size_t)blockIdx.x * world + rank, epoch);
        // ">= epoch" (wrap-around safe), not "== epoch": a peer that is already one launch ahead has
        // overwritten its flag with epoch + 1; its half for THIS epoch is intact until it passes the
        // handshake of that next launch, which needs this rank to get there too
        while ((int)(ld_acquire_sys(my_flags + (size_t)blockIdx.x * world + peer) - epoch) < 0) {
        }
    }
    __syncthreads();
    const double step = (double)(counter[0] + 1);
    const float bc1 = (float)(1.0 - pow((double)b1, step));
    const float bc2_sqrt = (float)sqrt(1.0 - pow((double)b2, step));
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)nb * blockDim.x) {
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {
            const float4 q = r == rank ? reinterpret_cast<const float4*>(grad_local)[i]
                                       : __ldcg(reinterpret_cast<const float4*>(peer_stage[r] + half) + i);
            g.x += q.x; g.y += q.y; g.z += q.z; g.w += q.w;
        }
        float gv[4] = {g.x, g.y, g.z, g.w};
        float4 pq = reinterpret_cast<float4*>(p)[i], mq = reinterpret_cast<float4*>(m)[i],
               vq = reinterpret_cast<float4*>(v)[i];
        float pv[4] = {pq.x, pq.y, pq.z, pq.w}, mv[4] = {mq.x, mq.y, mq.z, mq.w}, vv[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float gi = gv[k] * gscale;
            if (clip > 0.f && gi == gi) gi = fminf(fmaxf(gi, -clip), clip);
            if (wd != 0.f) gi = fmaf(wd, pv[k], gi);
            mv[k] = b1 * mv[k] + (1.f - b1) * gi;
            vv[k] = b2 * vv[k] + (1.f - b2) * gi * gi;
            const float denom = sqrtf(vv[k]) / bc2_sqrt + eps;
            pv[k] = pv[k] - (lr / bc1) * (mv[k] / denom);
        }
        reinterpret_cast<float4*>(p)[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
        reinterpret_cast<float4*>(m)[i] = make_float4(mv[0], mv[1], mv[2], mv[3]);
        reinterpret_cast<float4*>(v)[i] = make_float4(vv[0], vv[1], vv[2], vv[3]);
    }
    if (blockIdx.x == 0) {  // tail (n % 4 elements)
        for (int64_t i = (n4 << 2) + threadIdx.x; i < n; i += blockDim.x) {
            float gi = 0.f;
            for (int r = 0; r < world; ++r) gi += r == rank ? grad_local[i] : __ldcg(peer_stage[r] + half + i);
            gi *= gscale;
            if (clip > 0.f && gi == gi) gi = fminf(fmaxf(gi, -clip), clip);
            const float pi = p[i];
            if (wd != 0.f) gi = fmaf(wd, pi, gi);
            const float mi = b1 * m[i] + (1.f - b1) * gi, vi = b2 * v[i] + (1.f - b2) * gi * gi;
            m[i] = mi;
            v[i] = vi;
            p[i] = pi - (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == nb - 1) {  // last block of this launch: advance epoch and Adam step
            *ticket = 0u;
            epoch_ptr[0] = epoch;
            counter[0] += 1;
        }
    }
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_ce_loss(const float* logits, const int64_t* target, const float* class_weight,
                int64_t ignore_index, float* loss_out, float* d_logits, float* workspace,
                int64_t n_rows, int C, spg_stream_t stream) {
    if (n_rows <= 0 || C <= 0 || !logits || !target || !loss_out || !workspace)
        return SPG_E_BADARG;
    if (((uintptr_t)workspace & 7) != 0) return SPG_E_ALIGN;
    cudaStream_t s = (cudaStream_t)stream;
    cudaError_t e = cudaMemsetAsync(workspace, 0, 2 * sizeof(double), s);
    if (e != cudaSuccess) return (int)e;
    SPG_LAUNCH(K_CE_LOSS, s, ce_loss_kernel, (unsigned)ceil_div64(n_rows * 32, 256), 256, 0, logits,
               target, class_weight, ignore_index, (double*)workspace, d_logits, n_rows, C);
    int rc = launch_status();
    if (rc) return rc;
    const int64_t total = n_rows * C;
    SPG_LAUNCH(K_CE_LOSS_FINAL, s, ce_loss_final_kernel, (unsigned)ceil_div64(total, 256), 256, 0,
               (const double*)workspace, loss_out, d_logits, total);
    return launch_status();
}

int spg_clamp_adam(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                   float grad_clip, float grad_scale, int64_t step, spg_stream_t stream) {
    if (n < 0 || step < 1) return SPG_E_BADARG;
    if (n == 0) return SPG_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return SPG_E_BADARG;
    const double bc1 = 1.0 - pow((double)beta1, (double)step);
    const double bc2 = 1.0 - pow((double)beta2, (double)step);
    int64_t blocks = ceil_div64(n, 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    SPG_LAUNCH(K_CLAMP_ADAM, (cudaStream_t)stream, clamp_adam_kernel, (unsigned)blocks, 256, 0,
               param, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, grad_clip,
               grad_scale, (float)bc1, (float)sqrt(bc2));
    return launch_status();
}


int spg_clamp_adam_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                       float lr, float beta1, float beta2, float eps, float weight_decay,
                       float grad_clip, float grad_scale, int64_t* step_counter,
                       spg_stream_t stream) {
    if (n < 0 || !step_counter) return SPG_E_BADARG;
    if (n == 0) return SPG_OK;
    if (!param || !grad || !exp_avg || !exp_avg_sq) return SPG_E_BADARG;
    int64_t blocks = ceil_div64(n, 256);
    if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
    cudaStream_t s = (cudaStream_t)stream;
    SPG_LAUNCH(K_CLAMP_ADAM, s, clamp_adam_dev_kernel, (unsigned)blocks, 256, 0, param, grad, exp_avg,
               exp_avg_sq, n, lr, beta1, beta2, eps, weight_decay, grad_clip, grad_scale,
               (const long long*)step_counter);
    int rc = launch_status();
    if (rc) return rc;
    SPG_LAUNCH(K_CLAMP_ADAM, s, counter_increment_kernel, 1, 1, 0, (long long*)step_counter);
    return launch_status();
}


int spg_allreduce_flag_words(int world) { return 2 * kNumSMs * world + 8; }

int64_t spg_allreduce_stage_floats(int64_t n) { return n < 0 ? 0 : 2 * ((n + 3) / 4 * 4); }

int spg_allreduce_clamp_adam(const float* grad, float* const* peer_stage, uint32_t* const* peer_flags, int rank,
                             int world, float* param, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, float grad_clip,
                             float grad_scale, int64_t* step_counter, uint32_t* local_state, spg_stream_t stream) {
    if (n < 0 || world < 1 || world > 32 || rank < 0 || rank >= world) return SPG_E_BADARG;
    if (n == 0) return SPG_OK;
    if (!grad || !peer_stage || !peer_flags || !param || !exp_avg || !exp_avg_sq || !step_counter || !local_state)
        return SPG_E_BADARG;
    if (((uintptr_t)param | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq | (uintptr_t)grad) & 15) return SPG_E_ALIGN;
    int64_t blocks = ceil_div64(ceil_div64(n, 4), 512);
    if (blocks > 2 * kNumSMs) blocks = 2 * kNumSMs;
    if (blocks < 1) blocks = 1;
    SPG_LAUNCH(K_CLAMP_ADAM, (cudaStream_t)stream, allreduce_clamp_adam_kernel, (unsigned)blocks, 512, 0, grad,
               peer_stage, (unsigned* const*)peer_flags, rank, world, param, exp_avg, exp_avg_sq, n,
               (n + 3) / 4 * 4, lr, beta1, beta2, eps, weight_decay, grad_clip, grad_scale,
               (long long*)step_counter, (unsigned*)local_state, (unsigned*)local_state + 1);
    return launch_status();
}

}  // extern "C"
