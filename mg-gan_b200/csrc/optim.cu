// Fused gradient clipping + AdamW over a table of parameter tensors (reference:
// torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step called per tensor,
// mggan/model/train.py:131-135, :209-213, :656-658; optimiser set-up mggan/abstract_train.py:45-57).
//
// The table (<= 64 tensors per launch) travels in kernel-parameter space, so nothing is copied
// or allocated: pass 1 accumulates the global squared gradient norm into a device double,
// pass 2 applies  p <- p (1 - lr wd);  m, v updates;  p <- p - lr/bc1 * m / (sqrt(v)/sqrt(bc2) + eps)
// with the clip coefficient min(1, max_norm / (norm + 1e-6)) read from device memory (no host
// sync).  Per-tensor bias corrections are supplied by the host because AdamW's step count is
// per tensor (tensors without a gradient are skipped and keep their count); when the
// iteration is replayed as a CUDA graph they (and lr) are read from a device array the host refreshes.
#include "common.cuh"

#define MGGAN_TABLE_MAX 64

struct MgganTensorTable {
    float* p[MGGAN_TABLE_MAX];
    const float* g[MGGAN_TABLE_MAX];
    float* m[MGGAN_TABLE_MAX];
    float* v[MGGAN_TABLE_MAX];
    int n[MGGAN_TABLE_MAX];
    float bc1[MGGAN_TABLE_MAX];       // 1 - beta1^t
    float bc2_sqrt[MGGAN_TABLE_MAX];  // sqrt(1 - beta2^t)
    const float* dyn;                 // optional DEVICE array [lr, bc1[64], bc2_sqrt[64]] that overrides lr / bc1 / bc2_sqrt
                                      // (CUDA-graph replay: the values change every iteration, the launch does not)
};

namespace {

__global__ void __launch_bounds__(MGGAN_THREADS)
sqnorm_kernel(MgganTensorTable tb, double* __restrict__ out) {
    const int t = blockIdx.y;
    const float* g = tb.g[t];
    const int n = tb.n[t];
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float v = g[i];
        acc = fmaf(v, v, acc);
    }
    __shared__ float sred[MGGAN_THREADS / 32];
    acc = warp_sum(acc);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < MGGAN_THREADS / 32; ++w) s += (double)sred[w];
        if (s != 0.0) atomicAdd(out, s);
    }
}

__global__ void __launch_bounds__(MGGAN_THREADS)
adamw_kernel(MgganTensorTable tb, const double* __restrict__ sqnorm, float max_norm, float grad_scale, float lr,
             float beta1, float beta2, float eps, float wd) {
    const int t = blockIdx.y;
    float coef = grad_scale;
    if (max_norm > 0.f && sqnorm != nullptr) {
        float nrm = (float)sqrt(*sqnorm) * grad_scale;
        coef *= fminf(1.f, max_norm / (nrm + 1e-6f));
    }
    float* p = tb.p[t];
    const float* g = tb.g[t];
    float* m = tb.m[t];
    float* v = tb.v[t];
    const int n = tb.n[t];
    float bc1 = tb.bc1[t], bc2s = tb.bc2_sqrt[t];
    if (tb.dyn != nullptr) {
        lr = __ldg(tb.dyn);
        bc1 = __ldg(tb.dyn + 1 + t);
        bc2s = __ldg(tb.dyn + 1 + MGGAN_TABLE_MAX + t);
    }
    const float step = lr / bc1, decay = 1.f - lr * wd;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float gi = g[i] * coef;
        float mi = beta1 * m[i] + (1.f - beta1) * gi;
        float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
        m[i] = mi;
        v[i] = vi;
        float denom = sqrtf(vi) / bc2s + eps;
        p[i] = p[i] * decay - step * (mi / denom);
    }
}

// dst[t] <- src[t] (flat packing / unpacking for the gradient all-reduce)
__global__ void __launch_bounds__(MGGAN_THREADS)
multi_copy_kernel(MgganTensorTable tb) {
    const int t = blockIdx.y;
    float* dst = tb.p[t];
    const float* src = tb.g[t];
    const int n = tb.n[t];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[i];
}

int grid_x(const MgganTensorTable& tb, int count) {
    int mx = 1;
    for (int t = 0; t < count; ++t) mx = tb.n[t] > mx ? tb.n[t] : mx;
    int g = (mx + MGGAN_THREADS * 4 - 1) / (MGGAN_THREADS * 4);
    return g < 1 ? 1 : (g > 64 ? 64 : g);
}

}  // namespace

extern "C" int mggan_grad_sqnorm(const MgganTensorTable* table, int count, double* sqnorm_accum, cudaStream_t stream) {
    MGGAN_REQUIRE(count >= 0 && count <= MGGAN_TABLE_MAX, "mggan_grad_sqnorm: %d tensors (max %d per call)", count,
                  MGGAN_TABLE_MAX);
    if (count == 0) return MGGAN_OK;
    sqnorm_kernel<<<dim3(grid_x(*table, count), count), MGGAN_THREADS, 0, stream>>>(*table, sqnorm_accum);
    return mggan_check_launch("grad_sqnorm");
}

extern "C" int mggan_clip_adamw(const MgganTensorTable* table, int count, const double* sqnorm, float max_norm,
                                float grad_scale, float lr, float beta1, float beta2, float eps, float weight_decay,
                                cudaStream_t stream) {
    MGGAN_REQUIRE(count >= 0 && count <= MGGAN_TABLE_MAX, "mggan_clip_adamw: %d tensors (max %d per call)", count,
                  MGGAN_TABLE_MAX);
    if (count == 0) return MGGAN_OK;
    adamw_kernel<<<dim3(grid_x(*table, count), count), MGGAN_THREADS, 0, stream>>>(*table, sqnorm, max_norm, grad_scale,
                                                                                  lr, beta1, beta2, eps, weight_decay);
    return mggan_check_launch("clip_adamw");
}

extern "C" int mggan_multi_copy(const MgganTensorTable* table, int count, cudaStream_t stream) {
    MGGAN_REQUIRE(count >= 0 && count <= MGGAN_TABLE_MAX, "mggan_multi_copy: %d tensors (max %d per call)", count,
                  MGGAN_TABLE_MAX);
    if (count == 0) return MGGAN_OK;
    multi_copy_kernel<<<dim3(grid_x(*table, count), count), MGGAN_THREADS, 0, stream>>>(*table);
    return mggan_check_launch("multi_copy");
}
