// Data-side operators next to the training step (8a a17): the 4 x 33 x 33 crop features cut from
// scene images that stay resident in HBM, the growing-radius tube ("manifold") inside test behind Precision / Recall,
// and the scene-level min-over-samples ADE / FDE.  All three are small streaming / integer-decision kernels: HBM-bound
// (crop: 17,424 B written per agent, 3,267 B read) or latency-bound (metrics on evaluation sets); no tensor cores.
//
// Bit-exactness notes.  The crop reproduces the reference's float32 centre computation (multiply in fp32, truncate toward
// zero) and its exact -1 + u8 / 128 values.  The tube test reproduces numpy's float32 norm: (dx * dx + dy * dy) with
// separately rounded products (no FMA contraction), correctly rounded sqrt, compared with the float64 radius.
#include "common.cuh"

#define CROP 33
#define CROP_PIX (CROP * CROP)
#define CROP_MARGIN 16

// ---------------------------------------------------------------------------------------------------------- crop
// One CTA walks agents; a thread writes consecutive output floats (coalesced 128-byte warp stores), reading the three
// bytes of an RGB pixel from the atlas through the read-only path (neighbouring lanes read neighbouring pixels).
__global__ void __launch_bounds__(MGGAN_THREADS) scene_crop_kernel(
    const unsigned char* __restrict__ atlas, const long long* __restrict__ img_off, const int* __restrict__ img_wh,
    const float* __restrict__ img_scale, int n_images, const int* __restrict__ agent_img,
    const float* __restrict__ last_xy, int N, float* __restrict__ features) {
    for (int a = blockIdx.x; a < N; a += gridDim.x) {
        const int im = agent_img[a];
        const bool has = im >= 0 && im < n_images;
        int w = 0, h = 0, x0 = 0, y0 = 0;
        const unsigned char* pix = atlas;
        if (has) {
            w = img_wh[2 * im];
            h = img_wh[2 * im + 1];
            const float s = img_scale[im];
            // center_pixel_small = center_meter * scale (float32); astype(int) truncates toward zero
            x0 = (int)__fmul_rn(last_xy[2 * a], s) - CROP_MARGIN;
            y0 = (int)__fmul_rn(last_xy[2 * a + 1], s) - CROP_MARGIN;
            pix = atlas + img_off[im];
        }
        float* out = features + (size_t)a * 4 * CROP_PIX;
        for (int i = threadIdx.x; i < 4 * CROP_PIX; i += blockDim.x) {
            const int c = i / CROP_PIX;
            const int p = i - c * CROP_PIX;
            float v;
            if (c == 3) {
                v = (p == CROP_MARGIN * CROP + CROP_MARGIN) ? 1.f : 0.f;
            } else {
                const int yy = p / CROP;
                const int y = y0 + yy, x = x0 + (p - yy * CROP);
                unsigned int u = 0;                                   // PIL pads a crop outside the image with zeros
                if (x >= 0 && x < w && y >= 0 && y < h) u = __ldg(pix + ((size_t)y * w + x) * 3 + c);
                v = __fmaf_rn((float)u, 2.f / 256.f, -1.f);           // exact: u / 128 - 1
            }
            out[i] = v;
        }
    }
}

extern "C" int mggan_scene_crop(const unsigned char* atlas, const long long* img_off, const int* img_wh,
                                const float* img_scale, int n_images, const int* agent_img, const float* last_xy, int N,
                                float* features, cudaStream_t stream) {
    MGGAN_REQUIRE(N >= 0 && n_images >= 0, "mggan_scene_crop: negative size");
    if (N == 0) return MGGAN_OK;
    MGGAN_REQUIRE(agent_img && last_xy && features, "mggan_scene_crop: null pointer");
    MGGAN_REQUIRE(n_images == 0 || (atlas && img_off && img_wh && img_scale), "mggan_scene_crop: null image table");
    int grid = N < 148 * 8 ? N : 148 * 8;
    scene_crop_kernel<<<grid, MGGAN_THREADS, 0, stream>>>(atlas, img_off, img_wh, img_scale, n_images, agent_img, last_xy, N,
                                                        features);
    return mggan_check_launch("mggan_scene_crop");
}

// ---------------------------------------------------------------------------------------------------------- tube test
// One warp per test trajectory; lanes stride over the manifold samples of its descriptor, one ballot per time step.
__global__ void __launch_bounds__(MGGAN_THREADS) tube_inside_kernel(
    const float* __restrict__ traj, int T, const double* __restrict__ radius, const int* __restrict__ desc, int n_tests,
    const int* __restrict__ man_list, unsigned char* __restrict__ inside) {
    const int lane = threadIdx.x & 31;
    const int warps = (gridDim.x * blockDim.x) >> 5;
    for (int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; i < n_tests; i += warps) {
        const float* tp = traj + (size_t)desc[3 * i] * T * 2;
        const int first = desc[3 * i + 1], count = desc[3 * i + 2];
        bool all_t = true;
        for (int t = 0; t < T; ++t) {
            const float tx = tp[2 * t], ty = tp[2 * t + 1];
            const double r = radius[t];
            bool any = false;
            for (int j = lane; j < count; j += 32) {
                const float* mp = traj + ((size_t)man_list[first + j] * T + t) * 2;
                const float dx = __fsub_rn(mp[0], tx), dy = __fsub_rn(mp[1], ty);
                const float d = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
                any |= (double)d < r;                                 // NaN compares false, like numpy
            }
            if (!__any_sync(0xffffffffu, any)) {                      // uniform across the warp
                all_t = false;
                break;
            }
        }
        if (lane == 0) inside[i] = all_t ? 1 : 0;
    }
}

extern "C" int mggan_tube_inside(const float* traj, int T, const double* radius, const int* desc, int n_tests,
                                 const int* man_list, unsigned char* inside, cudaStream_t stream) {
    MGGAN_REQUIRE(T > 0 && n_tests >= 0, "mggan_tube_inside: T = %d, n_tests = %d", T, n_tests);
    if (n_tests == 0) return MGGAN_OK;
    MGGAN_REQUIRE(traj && radius && desc && man_list && inside, "mggan_tube_inside: null pointer");
    int blocks = (n_tests + 7) / 8;                                    // 8 warps per CTA
    if (blocks > 148 * 8) blocks = 148 * 8;
    tube_inside_kernel<<<blocks, MGGAN_THREADS, 0, stream>>>(traj, T, radius, desc, n_tests, man_list, inside);
    return mggan_check_launch("mggan_tube_inside");
}

// ---------------------------------------------------------------------------------------------------------- min ADE / FDE
// One CTA per scene.  A thread owns an agent and walks the K samples in order, so the per-agent running minimum of the
// final displacement (the "Mode" count) needs no exchange; per-sample error sums of the scene are accumulated in shared
// memory in double, then one thread takes the prefix minima over the samples: entry kk - 1 of a scene's row is the
// metric with the first kk predictions, which is what the k = 1 .. K sweep of scripts/evaluate.py asks for.
#define MINADE_MAX_K 64
__global__ void __launch_bounds__(MGGAN_THREADS) min_ade_fde_kernel(
    const float* __restrict__ preds, const float* __restrict__ gt, int T, int K, int n, const int* __restrict__ scene_off,
    int n_scenes, const float* __restrict__ scene_scale, float mode_thresh, double* __restrict__ ade,
    double* __restrict__ fde, int* __restrict__ mode) {
    __shared__ double s_ade[MINADE_MAX_K], s_fde[MINADE_MAX_K];
    __shared__ int s_mode[MINADE_MAX_K];
    for (int sc = blockIdx.x; sc < n_scenes; sc += gridDim.x) {
        for (int s = threadIdx.x; s < K; s += blockDim.x) {
            s_ade[s] = 0.0;
            s_fde[s] = 0.0;
            s_mode[s] = 0;
        }
        __syncthreads();
        const int a0 = scene_off[sc], a1 = scene_off[sc + 1];
        const float scale = scene_scale ? scene_scale[sc] : 1.f;
        for (int i = a0 + threadIdx.x; i < a1; i += blockDim.x) {
            float best = INFINITY;
            for (int s = 0; s < K; ++s) {
                float sum = 0.f, e = 0.f;
                for (int t = 0; t < T; ++t) {
                    const float* p = preds + (((size_t)t * K + s) * n + i) * 2;
                    const float* g = gt + ((size_t)t * n + i) * 2;
                    const float dx = p[0] * scale - g[0] * scale, dy = p[1] * scale - g[1] * scale;
                    e = sqrtf(dx * dx + dy * dy);
                    sum += e;
                }
                atomicAdd(&s_ade[s], (double)sum);
                atomicAdd(&s_fde[s], (double)e);
                best = fminf(best, e);
                if (best < mode_thresh) atomicAdd(&s_mode[s], 1);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ma = INFINITY, mf = INFINITY;
            for (int s = 0; s < K; ++s) {
                ma = fmin(ma, s_ade[s]);
                mf = fmin(mf, s_fde[s]);
                ade[(size_t)sc * K + s] = ma;
                fde[(size_t)sc * K + s] = mf;
                mode[(size_t)sc * K + s] = s_mode[s];
            }
        }
        __syncthreads();
    }
}

extern "C" int mggan_min_ade_fde(const float* preds, const float* gt, int T, int K, int n, const int* scene_off,
                                 int n_scenes, const float* scene_scale, float mode_thresh, double* ade, double* fde,
                                 int* mode, cudaStream_t stream) {
    MGGAN_REQUIRE(T > 0 && K > 0 && K <= MINADE_MAX_K && n >= 0 && n_scenes >= 0,
                  "mggan_min_ade_fde: T = %d, K = %d (1..%d), n = %d, scenes = %d", T, K, MINADE_MAX_K, n, n_scenes);
    if (n_scenes == 0) return MGGAN_OK;
    MGGAN_REQUIRE(preds && gt && scene_off && ade && fde && mode, "mggan_min_ade_fde: null pointer");
    int grid = n_scenes < 148 * 8 ? n_scenes : 148 * 8;
    min_ade_fde_kernel<<<grid, MGGAN_THREADS, 0, stream>>>(preds, gt, T, K, n, scene_off, n_scenes, scene_scale, mode_thresh,
                                                         ade, fde, mode);
    return mggan_check_launch("mggan_min_ade_fde");
}
