// The two steps either side of the training/inference path:
//
//  * cloud_build: the per-superpoint part of the reference's batch loader
//    (learning/spg.py:198-236 `load_superpoint`, :238-260 `augment_cloud`, stacked by
//    `loader` :146-166): resample every chosen superpoint to exactly L points, centre and scale
//    xyz, select the attribute columns, augment, and emit the [Nv, F, L] tensor PointNet reads.
//    The reference does this in numpy per superpoint from per-superpoint HDF5 datasets; here the
//    parsed points stay resident in HBM as one packed [rows, ldp] array and a CTA builds one cloud.
//    The arithmetic follows numpy's evaluation order (sequential fp32 column sums, fp32 IEEE
//    division) so that the result is bit-identical to the reference when the host supplies the
//    same sample indices.
//
//  * confusion_count: argmax + confusion-matrix accumulation of the evaluation loops
//    (learning/main.py:257-262,297-305; learning/metrics.py:16-18), integer, exact.
#include <float.h>

#include "common.cuh"

namespace spg {

// counter-based generator for the device-side sampling / jitter option (no numpy parity by
// construction: the reference draws from MT19937 on the host)
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float u01(uint64_t bits) {  // (0,1]
    return ((float)(bits >> 40) + 1.0f) * (1.0f / 16777216.0f);
}

struct CloudBuildArgs {
    const float* points;       // [rows, ldp] parsed point attributes of every resident superpoint
    int64_t ldp;
    const int64_t* sp_start;   // [Nv] first row of each selected superpoint
    const int32_t* sp_count;   // [Nv] its number of points (>= 1)
    const int32_t* sample_idx; // [Nv, L] source point of every output point, or null (device RNG)
    const int32_t* columns;    // [F] source column of each output attribute
    int F, L, normalize;
    const double* xform;       // [Nv, 9] row-major 3x3 applied to output attributes 0..2, or null
    const float* jitter;       // [Nv, L, F] additive noise (already clipped), or null
    float jitter_sigma, jitter_clip;  // used when jitter == null and jitter_sigma > 0
    uint64_t seed;
    float* clouds;             // [Nv, F, L]
    float* diameters;          // [Nv]
};

// One WARP per cloud (4 clouds per CTA): the kernel is a chain of dependent latencies (sample index
// -> point row -> sequential column sums -> normalise -> store), so throughput comes from the number
// of clouds in flight per SM (64 warps), not from the width of one cloud.  Only xyz is staged in
// shared memory (the sums need it in order); the other attributes are re-read from the row, which is
// in L1/L2 by then.  smem per warp: xyz[L][3] + 16 scalars; columns[F] once per CTA.
constexpr int kCloudWarps = 4;

__global__ void __launch_bounds__(kCloudWarps * 32) cloud_build_kernel(const CloudBuildArgs a, int64_t nv) {
    SPG_PDL_ENTRY();
    extern __shared__ float sm[];
    const int L = a.L, F = a.F;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int* cols = reinterpret_cast<int*>(sm);                 // [F]
    float* xyz = sm + F + (size_t)warp * (3 * L + 16);      // [L][3]
    float* red = xyz + 3 * L;                               // sum[3], min[3], max[3], denominator
    for (int f = threadIdx.x; f < F; f += blockDim.x) cols[f] = a.columns[f];
    __syncthreads();
    const int64_t i = (int64_t)blockIdx.x * kCloudWarps + warp;
    if (i >= nv) return;
    const int64_t start = a.sp_start[i];
    const int n = a.sp_count[i];
    auto src_row = [&](int j) -> int64_t {
        int r;
        if (a.sample_idx) {
            r = a.sample_idx[i * L + j];
        } else if (n == L || (n < L && j < n)) {
            r = j;  // kept as is / the original points come first (spg.py:209-214)
        } else {
            r = (int)(mix64(a.seed ^ mix64((uint64_t)i * 0x100000001B3ull + (uint64_t)j)) % (uint64_t)n);
        }
        return start + r;
    };
    // rows padded to a multiple of 4 floats (SuperpointStore does that) are fetched with 128-bit
    // loads: 1 + ldp/4 requests per point instead of 3 + F
    const bool vec = (a.ldp & 3) == 0 && a.ldp <= 16 && (reinterpret_cast<uintptr_t>(a.points) & 15) == 0;
    for (int j = lane; j < L; j += 32) {
        const float* p = a.points + src_row(j) * a.ldp;
        if (vec) {
            const float4 q = __ldg(reinterpret_cast<const float4*>(p));
            xyz[3 * j + 0] = q.x;
            xyz[3 * j + 1] = q.y;
            xyz[3 * j + 2] = q.z;
        } else {
            xyz[3 * j + 0] = __ldg(p + 0);
            xyz[3 * j + 1] = __ldg(p + 1);
            xyz[3 * j + 2] = __ldg(p + 2);
        }
    }
    __syncwarp();
    // numpy reduces a C-ordered [L,3] array over axis 0 row by row: plain sequential fp32 sums
    // (lanes 0-2); min and max are order-free (lanes 3-5, 6-8) and ride along in the same loop.
    if (lane < 9) {
        const int k = lane % 3, what = lane / 3;
        float v = xyz[k];
        for (int j = 1; j < L; ++j) {
            const float x = xyz[3 * j + k];
            v = what == 0 ? v + x : (what == 1 ? fminf(v, x) : fmaxf(v, x));
        }
        red[3 * what + k] = v;
    }
    __syncwarp();
    if (lane == 0) {
        float diam = 0.f;
        if (a.normalize) diam = fmaxf(fmaxf(red[6] - red[3], red[7] - red[4]), red[8] - red[5]);
        a.diameters[i] = diam;
        red[9] = a.normalize ? (float)((double)diam + 1e-10) : 1.f;
        red[0] = red[0] / (float)L;
        red[1] = red[1] / (float)L;
        red[2] = red[2] / (float)L;
    }
    __syncwarp();
    const float den = red[9];
    const bool norm = a.normalize;
    const double* M = a.xform ? a.xform + i * 9 : nullptr;
    auto noise = [&](int j, int f) -> float {
        if (a.jitter) return a.jitter[(i * L + j) * F + f];
        if (a.jitter_sigma > 0.f) {
            const uint64_t h = mix64(a.seed ^ mix64(((uint64_t)i * L + j) * 64 + f + 0x5bd1e995ull));
            const float g = sqrtf(-2.f * logf(u01(h))) * cospif(2.f * u01(mix64(h)));
            return fminf(fmaxf(a.jitter_sigma * g, -a.jitter_clip), a.jitter_clip);
        }
        return 0.f;
    };
    const bool noisy = a.jitter || a.jitter_sigma > 0.f;
    for (int j = lane; j < L; j += 32) {
        const float* p = a.points + src_row(j) * a.ldp;
        float c3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const float d = xyz[3 * j + k] - red[k];
            c3[k] = norm ? __fdiv_rn(d, den) : d;
        }
        float4 q0, q1, q2, q3;
        q0 = q1 = q2 = q3 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (vec) {
            const float4* p4 = reinterpret_cast<const float4*>(p);
            q0 = __ldg(p4);
            if (a.ldp > 4) q1 = __ldg(p4 + 1);
            if (a.ldp > 8) q2 = __ldg(p4 + 2);
            if (a.ldp > 12) q3 = __ldg(p4 + 3);
        }
        auto pick = [&](int c) -> float {  // register select, no local-memory indexing
            const float4 q = c < 8 ? (c < 4 ? q0 : q1) : (c < 12 ? q2 : q3);
            const int k = c & 3;
            return k < 2 ? (k == 0 ? q.x : q.y) : (k == 2 ? q.z : q.w);
        };
        float o3[3] = {0.f, 0.f, 0.f};
        for (int f = 0; f < F; ++f) {
            const int c = cols[f];
            float v = c < 3 ? c3[c] : (vec ? pick(c) : __ldg(p + c));
            if (M && f < 3) {
                o3[f] = v;  // written below, after the 3x3
                continue;
            }
            if (noisy) v += noise(j, f);
            a.clouds[(i * F + f) * L + j] = v;
        }
        if (M) {
            // P[:, :3] = P[:, :3] . M^T in double, rounded once (spg.py:254)
            const int nf = F < 3 ? F : 3;
            for (int f = 0; f < nf; ++f) {
                double acc = 0.0;
                for (int k = 0; k < nf; ++k) acc += (double)o3[k] * M[3 * f + k];
                float v = (float)acc;
                if (noisy) v += noise(j, f);
                a.clouds[(i * F + f) * L + j] = v;
            }
        }
    }
}

// One warp per superpoint: pred = first argmax of its logits; nodes with a label add their
// per-class point histogram to column `pred` of the confusion matrix.
__global__ void __launch_bounds__(256)
confusion_count_kernel(const float* __restrict__ logits, int64_t ldl,
                       const int64_t* __restrict__ label_mode,
                       const int64_t* __restrict__ label_vec, int64_t ldv,
                       unsigned long long* __restrict__ cm, unsigned long long* __restrict__ counters,
                       int64_t* __restrict__ pred_out, int64_t n, int C) {
    SPG_PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;
    float best = -FLT_MAX;
    int arg = 0x7fffffff;
    bool seen = false;
    for (int c = lane; c < C; c += 32) {
        const float v = logits[i * ldl + c];
        if (!seen || v > best) {  // strict: keeps the first maximum within the lane
            best = v;
            arg = c;
            seen = true;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (oa != 0x7fffffff && (arg == 0x7fffffff || ob > best || (ob == best && oa < arg))) {
            best = ob;
            arg = oa;
        }
    }
    if (lane == 0 && pred_out) pred_out[i] = arg;
    const int64_t t = label_mode[i];
    if (t == -100) return;  // no ground truth (main.py:447-452 filter_valid)
    for (int c = lane; c < C; c += 32) {
        const int64_t add = label_vec[i * ldv + c];
        if (add != 0) atomicAdd(cm + (int64_t)c * C + arg, (unsigned long long)add);
    }
    if (lane == 0) {
        atomicAdd(counters + 0, 1ull);
        if (t == arg) atomicAdd(counters + 1, 1ull);
    }
}

// ------------------------------------------------------------------ label up-sampling
// Step after the path (partition/provider.py:630-687): predictions live on
// superpoints of a PRUNED cloud; the full cloud gets them by (a) scattering every superpoint's label
// to its member points and (b) exact 1-nearest-neighbour transfer from the pruned to the full cloud.
// The reference does (b) with scikit-learn's kd-tree in float64; here a thread owns a query point,
// the reference points stream through shared memory as doubles, and the squared distance is formed
// in float64 exactly as sklearn does ((x-y) exact for float32 inputs), so the neighbour index is the
// same whenever the nearest neighbour is unique; ties go to the lowest index.
constexpr int kNnTile = 1024;

__global__ void __launch_bounds__(256)
nn1_kernel(const float* __restrict__ ref, int64_t n_ref, const float* __restrict__ qry, int64_t n_q,
           const int64_t* __restrict__ labels_ref, int64_t* __restrict__ labels_out,
           int32_t* __restrict__ idx_out) {
    SPG_PDL_ENTRY();
    __shared__ double sx[kNnTile], sy[kNnTile], sz[kNnTile];
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double qx = 0.0, qy = 0.0, qz = 0.0;
    if (q < n_q) {
        qx = (double)qry[q * 3];
        qy = (double)qry[q * 3 + 1];
        qz = (double)qry[q * 3 + 2];
    }
    double best = 1.0e300;
    int64_t arg = 0;
    for (int64_t base = 0; base < n_ref; base += kNnTile) {
        const int n = (int)min((int64_t)kNnTile, n_ref - base);
        __syncthreads();
        for (int i = threadIdx.x; i < n; i += blockDim.x) {
            sx[i] = (double)ref[(base + i) * 3];
            sy[i] = (double)ref[(base + i) * 3 + 1];
            sz[i] = (double)ref[(base + i) * 3 + 2];
        }
        __syncthreads();
#pragma unroll 4
        for (int i = 0; i < n; ++i) {
            const double dx = qx - sx[i], dy = qy - sy[i], dz = qz - sz[i];
            const double d = dx * dx + dy * dy + dz * dz;
            if (d < best) {
                best = d;
                arg = base + i;
            }
        }
    }
    if (q < n_q) {
        if (idx_out) idx_out[q] = (int32_t)arg;
        if (labels_out) labels_out[q] = labels_ref[arg];
    }
}

// labels_full[point_ids[j]] = labels_red[component of j]; comp_ptr is the CSR over the concatenated
// member lists (one warp-strided pass; later components overwrite earlier ones as the Python loop does
// — members are disjoint in every SPG file).
__global__ void __launch_bounds__(256)
labels_to_points_kernel(const int64_t* __restrict__ labels_red, const int64_t* __restrict__ comp_ptr,
                        const int64_t* __restrict__ point_ids, int64_t n_comp,
                        uint8_t* __restrict__ labels_full, int64_t n_ver) {
    SPG_PDL_ENTRY();
    const int lane = threadIdx.x & 31;
    const int64_t c = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (c >= n_comp) return;
    const uint8_t lab = (uint8_t)labels_red[c];
    for (int64_t j = comp_ptr[c] + lane; j < comp_ptr[c + 1]; j += 32) {
        const int64_t v = point_ids[j];
        if (v >= 0 && v < n_ver) labels_full[v] = lab;
    }
}

}  // namespace spg

using namespace spg;

extern "C" {

int spg_cloud_build(const float* points, int64_t ldp, const int64_t* sp_start,
                    const int32_t* sp_count, const int32_t* sample_idx, const int32_t* columns,
                    int n_attribs, int n_points, int normalize, const double* xform,
                    const float* jitter, float jitter_sigma, float jitter_clip, int64_t seed,
                    float* clouds, float* diameters, int64_t n_clouds, spg_stream_t stream) {
    if (n_clouds < 0 || n_attribs <= 0 || n_points <= 0 || ldp < 3) return SPG_E_BADARG;
    if (n_clouds == 0) return SPG_OK;
    if (!points || !sp_start || !sp_count || !columns || !clouds || !diameters) return SPG_E_BADARG;
    const size_t smem = sizeof(float) * ((size_t)n_attribs + kCloudWarps * (3 * (size_t)n_points + 16));
    if (smem > 200 * 1024 || n_clouds > 0x7fffffffll) return SPG_E_UNSUPPORTED;
    CloudBuildArgs a;
    a.points = points;
    a.ldp = ldp;
    a.sp_start = sp_start;
    a.sp_count = sp_count;
    a.sample_idx = sample_idx;
    a.columns = columns;
    a.F = n_attribs;
    a.L = n_points;
    a.normalize = normalize;
    a.xform = xform;
    a.jitter = jitter;
    a.jitter_sigma = jitter_sigma;
    a.jitter_clip = jitter_clip;
    a.seed = (uint64_t)seed;
    a.clouds = clouds;
    a.diameters = diameters;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(cloud_build_kernel,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
    }
    SPG_LAUNCH(K_CLOUD_BUILD, (cudaStream_t)stream, cloud_build_kernel,
               (unsigned)ceil_div64(n_clouds, kCloudWarps), kCloudWarps * 32, smem, a, n_clouds);
    return launch_status();
}

int spg_confusion_count(const float* logits, int64_t ld_logits, const int64_t* label_mode,
                        const int64_t* label_vec, int64_t ld_vec, int64_t* confusion,
                        int64_t* counters, int64_t* pred_out, int64_t n_nodes, int n_classes,
                        spg_stream_t stream) {
    if (n_nodes < 0 || n_classes <= 0) return SPG_E_BADARG;
    if (n_nodes == 0) return SPG_OK;
    if (!logits || !label_mode || !label_vec || !confusion || !counters) return SPG_E_BADARG;
    const int64_t blocks = ceil_div64(n_nodes, 8);
    SPG_LAUNCH(K_CONFUSION, (cudaStream_t)stream, confusion_count_kernel, (unsigned)blocks, 256, 0,
               logits, ld_logits, label_mode, label_vec, ld_vec,
               reinterpret_cast<unsigned long long*>(confusion),
               reinterpret_cast<unsigned long long*>(counters), pred_out, n_nodes, n_classes);
    return launch_status();
}

int spg_nn1_interpolate(const float* xyz_ref, int64_t n_ref, const float* xyz_query, int64_t n_query,
                        const int64_t* labels_ref, int64_t* labels_out, int32_t* nn_index_out,
                        spg_stream_t stream) {
    if (n_ref < 0 || n_query < 0) return SPG_E_BADARG;
    if (n_query == 0) return SPG_OK;
    if (n_ref == 0 || !xyz_ref || !xyz_query || (!labels_out && !nn_index_out)) return SPG_E_BADARG;
    if (labels_out && !labels_ref) return SPG_E_BADARG;
    if (n_ref >= (1ll << 31)) return SPG_E_UNSUPPORTED;
    SPG_LAUNCH(K_CONFUSION, (cudaStream_t)stream, nn1_kernel, (unsigned)ceil_div64(n_query, 256), 256, 0,
               xyz_ref, n_ref, xyz_query, n_query, labels_ref, labels_out, nn_index_out);
    return launch_status();
}

int spg_labels_to_points(const int64_t* labels_red, const int64_t* comp_ptr, const int64_t* point_ids,
                         int64_t n_components, uint8_t* labels_full, int64_t n_ver, spg_stream_t stream) {
    if (n_components < 0 || n_ver < 0) return SPG_E_BADARG;
    if (n_ver == 0) return SPG_OK;
    if (!labels_full) return SPG_E_BADARG;
    cudaError_t e = cudaMemsetAsync(labels_full, 0, (size_t)n_ver, (cudaStream_t)stream);
    if (e != cudaSuccess) return (int)e;
    if (n_components == 0) return SPG_OK;
    if (!labels_red || !comp_ptr || !point_ids) return SPG_E_BADARG;
    SPG_LAUNCH(K_CONFUSION, (cudaStream_t)stream, labels_to_points_kernel,
               (unsigned)ceil_div64(n_components * 32, 256), 256, 0, labels_red, comp_ptr, point_ids,
               n_components, labels_full, n_ver);
    return launch_status();
}

}  // extern "C"
