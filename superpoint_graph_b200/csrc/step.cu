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

__global__ void counter_increment_kernel(long long* counter) { counter[0] += 1; }

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

}  // extern "C"
