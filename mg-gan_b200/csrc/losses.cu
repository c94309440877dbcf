// GAN-step losses and their gradients (reference: mggan/model/train.py:55-113 generator step,
// :148-200 discriminator step, :626-639 PM-Network "ml" target; objective lambdas
// mggan/abstract_train.py:61-85 for gan_obj = "NS": BCE with a scalar smoothed label).
//
// Every kernel produces the loss value and the gradient w.r.t. its inputs in one pass; the
// reference's Python loops over scenes (train.py:67-71) and generators (:94-96, :109-110) and
// its .item() syncs disappear.  Loss scalars are accumulated with one atomicAdd per CTA.
#include "common.cuh"

namespace {

__device__ __forceinline__ void block_add(float v, float* dst) {
    __shared__ float sred[MGGAN_THREADS / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sred[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += sred[w];
        atomicAdd(dst, s);
    }
}

// ---- scene-level min-over-samples L2 (train.py:57-75) ---------------------------------------
// abs (T, k, n, 2), gt (T, n, 2); scene ranges are the *un-adjusted* seq_start_end clipped to n
// (reference quirk).  One CTA per scene, one warp per sample.  squared: l2_loss_type "mse"
// (train.py:62-63: the per-step distances are squared before the sum over time).
__global__ void __launch_bounds__(MGGAN_THREADS)
l2_scene_min_kernel(const float* __restrict__ abs_, const float* __restrict__ gt, int T, int k, int n,
                    const int* __restrict__ scene_off, int n_scenes, float inv_norm, int squared,
                    float* __restrict__ loss, int* __restrict__ best, float* __restrict__ d_abs) {
    extern __shared__ float stot[];      // [k]
    const int sc = blockIdx.x;
    const int a = min(scene_off[sc], n), e = min(scene_off[sc + 1], n);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s = warp; s < k; s += MGGAN_THREADS / 32) {
        float acc = 0.f;
        for (int i = a + lane; i < e; i += 32) {
            float d = 0.f;
            for (int t = 0; t < T; ++t) {
                float2 p = __ldg(reinterpret_cast<const float2*>(abs_) + ((size_t)t * k + s) * n + i);
                float2 q = __ldg(reinterpret_cast<const float2*>(gt) + (size_t)t * n + i);
                float dx = p.x - q.x, dy = p.y - q.y;
                const float sq = dx * dx + dy * dy;
                d += squared ? sq : sqrtf(sq);
            }
            acc += d;
        }
        acc = warp_sum(acc);
        if (lane == 0) stot[s] = acc;
    }
    __syncthreads();
    __shared__ int sbest;
    if (threadIdx.x == 0) {
        int b = 0;
        float m = stot[0];
        for (int s = 1; s < k; ++s)
            if (stot[s] < m) { m = stot[s]; b = s; }
        sbest = b;
        best[sc] = b;
        atomicAdd(loss, m * inv_norm);
    }
    __syncthreads();
    if (d_abs == nullptr) return;
    const int b = sbest;
    for (int q = threadIdx.x; q < (e - a) * T; q += MGGAN_THREADS) {
        int t = q / (e - a), i = a + q - t * (e - a);
        size_t o = ((size_t)t * k + b) * n + i;
        float2 p = __ldg(reinterpret_cast<const float2*>(abs_) + o);
        float2 g = __ldg(reinterpret_cast<const float2*>(gt) + (size_t)t * n + i);
        float dx = p.x - g.x, dy = p.y - g.y;
        float nr = sqrtf(dx * dx + dy * dy);
        float sc_ = squared ? 2.f * inv_norm : (nr > 0.f ? inv_norm / nr : 0.f);     // d/dp of |p - g|^2 or |p - g|
        reinterpret_cast<float2*>(d_abs)[o] = make_float2(dx * sc_, dy * sc_);
    }
}

// ---- BCE against a scalar label, optional 1/count(generator) weights (train.py:92-97) --------
__global__ void __launch_bounds__(MGGAN_THREADS)
bce_kernel(const float* __restrict__ p, int n, float label, const long long* __restrict__ gen_idx,
           const int* __restrict__ counts, float inv_denom, float* __restrict__ loss, float* __restrict__ dp) {
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float pi = p[i];
        float w = inv_denom;
        if (gen_idx != nullptr) w /= (float)counts[gen_idx[i]];
        float l = -(label * fmaxf(logf(pi), -100.f) + (1.f - label) * fmaxf(log1pf(-pi), -100.f));
        acc = fmaf(w, l, acc);
        if (dp != nullptr) dp[i] = w * (pi - label) / fmaxf((1.f - pi) * pi, 1e-12f);
    }
    block_add(acc, loss);
}

// ---- least-squares GAN objective (gan_obj = "LS": nn.MSELoss(reduction="none"), abstract_train.py:72-75) ----
__global__ void __launch_bounds__(MGGAN_THREADS)
mse_kernel(const float* __restrict__ p, int n, float label, const long long* __restrict__ gen_idx,
           const int* __restrict__ counts, float inv_denom, float* __restrict__ loss, float* __restrict__ dp) {
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float w = inv_denom;
        if (gen_idx != nullptr) w /= (float)counts[gen_idx[i]];
        const float d = p[i] - label;
        acc = fmaf(w, d * d, acc);
        if (dp != nullptr) dp[i] = 2.f * w * d;
    }
    block_add(acc, loss);
}

// ---- cross-entropy over generators, optional 1/count weights (train.py:105-113, :184) --------
__global__ void __launch_bounds__(MGGAN_THREADS)
ce_kernel(const float* __restrict__ logits, int n, int G, const long long* __restrict__ target,
          const int* __restrict__ counts, float inv_denom, float* __restrict__ loss, float* __restrict__ dlogits) {
    float acc = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* z = logits + (size_t)i * G;
        int tg = (int)target[i];
        float mx = -INFINITY;
        for (int g = 0; g < G; ++g) mx = fmaxf(mx, z[g]);
        float sum = 0.f;
        for (int g = 0; g < G; ++g) sum += expf(z[g] - mx);
        float lse = mx + logf(sum);
        float w = inv_denom;
        if (counts != nullptr) w /= (float)counts[tg];
        acc = fmaf(w, lse - z[tg], acc);
        if (dlogits != nullptr)
            for (int g = 0; g < G; ++g) dlogits[(size_t)i * G + g] = w * (expf(z[g] - lse) - (g == tg ? 1.f : 0.f));
    }
    block_add(acc, loss);
}

// ---- PM-Network "ml" target (train.py:626-639) -------------------------------------------------
// abs_all (T, ks, G, n, 2) no-grad predictions of every generator; gt (T, n, 2); logits (n, G).
__global__ void __launch_bounds__(MGGAN_THREADS)
pm_ml_kernel(const float* __restrict__ abs_all, const float* __restrict__ gt, int T, int ks, int G, int n,
             const float* __restrict__ logits, float sigma, float weight, float inv_n, float* __restrict__ loss,
             float* __restrict__ dlogits, float* __restrict__ target_out) {
    constexpr int GMAX = 32;
    float acc = 0.f;
    const float inv2s2 = 1.f / (2.f * sigma * sigma);
    const float cst = -logf(sigma) - 0.9189385332046727f;      // -log(sigma) - log(sqrt(2 pi))
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float lp[GMAX];
#pragma unroll
        for (int g = 0; g < GMAX; ++g) {
            if (g >= G) { lp[g] = -INFINITY; continue; }
            float a = 0.f;
            for (int s = 0; s < ks; ++s)
                for (int t = 0; t < T; ++t) {
                    float2 p = __ldg(reinterpret_cast<const float2*>(abs_all) + (((size_t)t * ks + s) * G + g) * n + i);
                    float2 q = __ldg(reinterpret_cast<const float2*>(gt) + (size_t)t * n + i);
                    float dx = p.x - q.x, dy = p.y - q.y;
                    a += (cst - dx * dx * inv2s2) + (cst - dy * dy * inv2s2);
                }
            lp[g] = a / (float)ks;
        }
        float mx = -INFINITY, mz = -INFINITY;
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
            if (g < G) { mx = fmaxf(mx, lp[g]); mz = fmaxf(mz, logits[(size_t)i * G + g]); }
        float st = 0.f, sz = 0.f;
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
            if (g < G) { st += expf(lp[g] - mx); sz += expf(logits[(size_t)i * G + g] - mz); }
        float lsz = mz + logf(sz);
        float li = 0.f;
#pragma unroll
        for (int g = 0; g < GMAX; ++g)
            if (g < G) {
                float tau = expf(lp[g] - mx) / st;
                float z = logits[(size_t)i * G + g];
                li -= tau * (z - lsz);
                if (dlogits != nullptr) dlogits[(size_t)i * G + g] = weight * (expf(z - lsz) - tau) * inv_n;
                if (target_out != nullptr) target_out[(size_t)i * G + g] = tau;
            }
        acc += li * inv_n;
    }
    block_add(acc, loss);
}

}  // namespace

extern "C" int mggan_l2_scene_min(const float* abs_, const float* gt, int T, int k, int n, const int* scene_off,
                                  int n_scenes, float inv_norm, int squared, float* loss, int* best, float* d_abs,
                                  cudaStream_t stream) {
    MGGAN_REQUIRE(k >= 1 && T >= 1, "mggan_l2_scene_min: bad arguments");
    if (n_scenes <= 0) return MGGAN_OK;
    l2_scene_min_kernel<<<n_scenes, MGGAN_THREADS, sizeof(float) * k, stream>>>(abs_, gt, T, k, n, scene_off, n_scenes,
                                                                                inv_norm, squared, loss, best, d_abs);
    return mggan_check_launch("l2_scene_min");
}

extern "C" int mggan_bce_scalar_label(const float* p, int n, float label, const long long* gen_idx, const int* counts,
                                      float inv_denom, float* loss, float* dp, cudaStream_t stream) {
    if (n <= 0) return MGGAN_OK;
    int grid = (n + MGGAN_THREADS - 1) / MGGAN_THREADS;
    if (grid > 148 * 4) grid = 148 * 4;
    bce_kernel<<<grid, MGGAN_THREADS, 0, stream>>>(p, n, label, gen_idx, counts, inv_denom, loss, dp);
    return mggan_check_launch("bce_scalar_label");
}

extern "C" int mggan_mse_scalar_label(const float* p, int n, float label, const long long* gen_idx, const int* counts,
                                      float inv_denom, float* loss, float* dp, cudaStream_t stream) {
    if (n <= 0) return MGGAN_OK;
    int grid = (n + MGGAN_THREADS - 1) / MGGAN_THREADS;
    if (grid > 148 * 4) grid = 148 * 4;
    mse_kernel<<<grid, MGGAN_THREADS, 0, stream>>>(p, n, label, gen_idx, counts, inv_denom, loss, dp);
    return mggan_check_launch("mse_scalar_label");
}

extern "C" int mggan_ce_generators(const float* logits, int n, int G, const long long* target, const int* counts,
                                   float inv_denom, float* loss, float* dlogits, cudaStream_t stream) {
    MGGAN_REQUIRE(G >= 1, "mggan_ce_generators: bad G");
    if (n <= 0) return MGGAN_OK;
    int grid = (n + MGGAN_THREADS - 1) / MGGAN_THREADS;
    if (grid > 148 * 4) grid = 148 * 4;
    ce_kernel<<<grid, MGGAN_THREADS, 0, stream>>>(logits, n, G, target, counts, inv_denom, loss, dlogits);
    return mggan_check_launch("ce_generators");
}

extern "C" int mggan_pm_ml_loss(const float* abs_all, const float* gt, int T, int ks, int G, int n, const float* logits,
                                float sigma, float weight, float inv_n, float* loss, float* dlogits, float* target_out,
                                cudaStream_t stream) {
    MGGAN_REQUIRE(G >= 1 && G <= 32 && ks >= 1, "mggan_pm_ml_loss: num_gens %d not in [1, 32]", G);
    if (n <= 0) return MGGAN_OK;
    int grid = (n + MGGAN_THREADS - 1) / MGGAN_THREADS;
    if (grid > 148 * 4) grid = 148 * 4;
    pm_ml_kernel<<<grid, MGGAN_THREADS, 0, stream>>>(abs_all, gt, T, ks, G, n, logits, sigma, weight, inv_n, loss,
                                                     dlogits, target_out);
    return mggan_check_launch("pm_ml_loss");
}
