// Social-Ways attention pooling over the agents of one scene (reference:
// mggan/model/modules/social.py:7-123 — SocialFeatures/DCA_MTX/BearingMTX :67-104,
// EmbedSocialFeatures :33-48, AttentionPooling :14-30).
//
// For an ordered in-scene pair (i, j):  f = (|dp|, cos bearing, distance of closest approach),
// a1 = ReLU(W1 f + b1) (32), a2 = ReLU(W2 a1 + b2) (64), e = W3 a2 + b3, sigma_ij = e . (W h_j + b).
// The last embedding layer and the attention projection are linear in a2 and in h_j, so the
// host folds them into per-agent vectors  Us[j] = (u_j, s_j),  u_j = W3^T (W h_j + b) (64),
// s_j = b3 . (W h_j + b):  sigma_ij = a2 . u_j + s_j.  sigma_ii = -1000, att = softmax_j,
// S_i = sum_j att_ij h_j; single-agent scenes give 0 (social.py:19-20).
//
// The reference evaluates the pair MLP on all rows^2 pairs and then loops over agents in
// Python; here a CTA owns a scene (only the n_s^2 in-scene pairs exist), a warp owns agent i,
// lanes own neighbours j.  The backward recomputes the pair MLP and turns the per-pair outer
// products of the weight gradients into shared-memory tile products.
#include "common.cuh"
#include <math.h>

namespace {

constexpr int F1 = 32, F2 = 64;      // hidden widths of the pair MLP (social.py:38-44)
constexpr int LDU = F2 + 1;          // Us row: 64 + 1
constexpr int NMAX = 64;             // scenes up to this size are staged in shared memory
constexpr int LDW2 = F1 + 4;         // 36
constexpr int PT = 128;              // pairs per backward tile
constexpr int LDA2 = F2 + 4;         // 68
constexpr int LDA1 = F1 + 4;         // 36

__device__ __forceinline__ void pair_features(const float4 xi, const float4 xj, float& f1, float& f2, float& f3) {
    float dpx = xi.x - xj.x, dpy = xi.y - xj.y, dvx = xi.z - xj.z, dvy = xi.w - xj.w;
    f1 = sqrtf(dpx * dpx + dpy * dpy);
    float vn = sqrtf(xi.z * xi.z + xi.w * xi.w);
    f2 = (dpx * xi.z + dpy * xi.w) / (f1 * vn + 1e-6f);
    float ttca = -(dpx * dvx + dpy * dvy) / (dvx * dvx + dvy * dvy + 1e-6f);
    float cx = dpx + ttca * dvx, cy = dpy + ttca * dvy;
    f3 = sqrtf(cx * cx + cy * cy);
}

// a1 = ReLU(W1 f + b1); sW1 rows are (w0, w1, w2, b).
__device__ __forceinline__ void pair_layer1(const float* __restrict__ sW1, float f1, float f2, float f3, float (&a1)[F1]) {
#pragma unroll
    for (int k = 0; k < F1; ++k) {
        float4 w = ld4(sW1 + 4 * k);
        a1[k] = fmaxf(fmaf(w.x, f1, fmaf(w.y, f2, fmaf(w.z, f3, w.w))), 0.f);
    }
}

__device__ __forceinline__ float pair_layer2_unit(const float* __restrict__ sW2, const float* __restrict__ sb2, int c,
                                                   const float (&a1)[F1]) {
    float s = sb2[c];
#pragma unroll
    for (int k = 0; k < F1; k += 4) {
        float4 w = ld4(sW2 + c * LDW2 + k);
        s = fmaf(w.x, a1[k], fmaf(w.y, a1[k + 1], fmaf(w.z, a1[k + 2], fmaf(w.w, a1[k + 3], s))));
    }
    return s;
}

__device__ __forceinline__ void stage_pair_weights(float* sW1, float* sW2, float* sb2, const float* W1, const float* b1,
                                                   const float* W2, const float* b2) {
    for (int i = threadIdx.x; i < F1; i += blockDim.x) {
        sW1[4 * i] = __ldg(W1 + 3 * i); sW1[4 * i + 1] = __ldg(W1 + 3 * i + 1);
        sW1[4 * i + 2] = __ldg(W1 + 3 * i + 2); sW1[4 * i + 3] = __ldg(b1 + i);
    }
    stage_matrix(sW2, LDW2, W2, F2, F1);
    for (int i = threadIdx.x; i < F2; i += blockDim.x) sb2[i] = __ldg(b2 + i);
}

template <int HD>
__global__ void __launch_bounds__(MGGAN_THREADS)
social_fwd_kernel(const float* __restrict__ x4, const float* __restrict__ h, const float* __restrict__ Us,
                  const int* __restrict__ scene_off, const int* __restrict__ pair_off, int n_scenes,
                  const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                  const float* __restrict__ b2, float* __restrict__ S, float* __restrict__ att) {
    constexpr int LDHS = HD + 1;
    extern __shared__ __align__(16) float smem[];
    float* sW2 = smem;                       // [F2][LDW2]
    float* sW1 = sW2 + F2 * LDW2;            // [F1][4]
    float* sb2 = sW1 + F1 * 4;               // [F2]
    float* sX = sb2 + F2;                    // [NMAX][4]
    float* sU = sX + NMAX * 4;               // [NMAX][LDU]
    float* sHh = sU + NMAX * LDU;            // [NMAX][LDHS]
    stage_pair_weights(sW1, sW2, sb2, W1, b1, W2, b2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int sc = blockIdx.x; sc < n_scenes; sc += gridDim.x) {
        const int a = scene_off[sc], n = scene_off[sc + 1] - a;
        if (n <= 1) {
            for (int i = threadIdx.x; i < n * HD; i += MGGAN_THREADS) S[(size_t)a * HD + i] = 0.f;
            continue;
        }
        const bool staged = n <= NMAX;
        __syncthreads();
        if (staged) {
            for (int i = threadIdx.x; i < n * 4; i += MGGAN_THREADS) sX[i] = __ldg(x4 + (size_t)a * 4 + i);
            for (int i = threadIdx.x; i < n * LDU; i += MGGAN_THREADS) sU[i] = __ldg(Us + (size_t)a * LDU + i);
            for (int i = threadIdx.x; i < n * HD; i += MGGAN_THREADS)
                sHh[(i / HD) * LDHS + (i % HD)] = __ldg(h + (size_t)a * HD + i);
        }
        __syncthreads();
        const float* pX = staged ? sX : x4 + (size_t)a * 4;
        const float* pU = staged ? sU : Us + (size_t)a * LDU;
        const float* pH = staged ? sHh : h + (size_t)a * HD;
        const int ldh = staged ? LDHS : HD;
        const size_t poff = (size_t)pair_off[sc];
        for (int il = warp; il < n; il += MGGAN_THREADS / 32) {
            const float4 xi = ld4(pX + 4 * il);
            float* arow = att + poff + (size_t)il * n;
            float mx = -INFINITY;
            for (int jl = lane; jl < n; jl += 32) {
                float sigma = -1000.f;
                if (jl != il) {
                    float f1, f2, f3, a1[F1];
                    pair_features(xi, ld4(pX + 4 * jl), f1, f2, f3);
                    pair_layer1(sW1, f1, f2, f3, a1);
                    const float* uj = pU + (size_t)jl * LDU;
                    sigma = uj[F2];
#pragma unroll 4
                    for (int c = 0; c < F2; ++c) {
                        float s = pair_layer2_unit(sW2, sb2, c, a1);
                        sigma = fmaf(fmaxf(s, 0.f), uj[c], sigma);
                    }
                }
                arow[jl] = sigma;
                mx = fmaxf(mx, sigma);
            }
            mx = warp_max(mx);
            float sum = 0.f;
            for (int jl = lane; jl < n; jl += 32) {
                float e = __expf(arow[jl] - mx);
                arow[jl] = e;
                sum += e;
            }
            sum = warp_sum(sum);
            const float inv = 1.f / sum;
            for (int jl = lane; jl < n; jl += 32) arow[jl] *= inv;
            __syncwarp();
            for (int k = lane; k < HD; k += 32) {
                float acc = 0.f;
                for (int jl = 0; jl < n; ++jl) acc = fmaf(arow[jl], pH[(size_t)jl * ldh + k], acc);
                S[(size_t)(a + il) * HD + k] = acc;
            }
        }
    }
}

// Backward.  Phase 1 (per scene): softmax backward -> d sigma (scratch `dsig`) and dh_j = sum_i att_ij dS_i as a
// small matrix product (no atomics).  Phase 2: the pair MLP backward over tiles of PT pairs in j-major order
// (pair p -> j = p / n, i = p % n), every dense piece as a register-blocked shared-memory tile product:
//     S  = A1 W2^T + b2        (PT x 64, K = 32)     recomputed pre-activation of layer 2
//     D2 = [S > 0] ds u_j ,  dU_j += ds relu(S)      (run-length accumulated per thread, flushed with atomics)
//     dA1 = D2 W2 [A1 > 0]     (PT x 32, K = 64)
//     dW2 += D2^T A1, db2 += colsum D2, dW1 += dA1^T F, db1 += colsum dA1   (kept in registers across tiles)
template <int HD>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
social_bwd_kernel(const float* __restrict__ x4, const float* __restrict__ h, const float* __restrict__ Us,
                  const int* __restrict__ scene_off, const int* __restrict__ pair_off, int n_scenes,
                  const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                  const float* __restrict__ b2, const float* __restrict__ att, const float* __restrict__ dS,
                  float* __restrict__ dsig, float* __restrict__ dh, float* __restrict__ dUs, float* __restrict__ dW1,
                  float* __restrict__ db1, float* __restrict__ dW2, float* __restrict__ db2) {
    constexpr int LDHS = HD + 1;
    extern __shared__ __align__(16) float smem[];
    float* sW2 = smem;                       // [F2][LDW2]
    float* sW1 = sW2 + F2 * LDW2;            // [F1][4]
    float* sb2 = sW1 + F1 * 4;               // [F2]
    float* sX = sb2 + F2;                    // [NMAX][4]
    float* sU = sX + NMAX * 4;               // [NMAX][LDU]
    float* sA1 = sU + ((NMAX * LDU + 3) & ~3);   // [PT][LDA1]          (phase 1: sHh [NMAX][LDHS])
    float* sD2 = sA1 + PT * LDA1;            // [PT][LDA2]          (phase 1: sdS [NMAX][LDHS])
    float* sDA1 = sD2 + PT * LDA2;           // [PT][LDA1]
    float* sF = sDA1 + PT * LDA1;            // [PT][4]
    float* sDs = sF + PT * 4;                // [PT] d sigma of the pair
    int* sJ = reinterpret_cast<int*>(sDs + PT);   // [PT] local neighbour index j of the pair (-1: padding)
    float* sHh = sA1;
    float* sdS = sD2;
    stage_pair_weights(sW1, sW2, sb2, W1, b1, W2, b2);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // layer-2 tile product: warp pair -> 32 rows, thread -> rows rl + 4 i (i < 8), columns u + 16 jj (jj < 4)
    const int u = (warp & 1) * 8 + (lane & 7);
    const int rl = (warp >> 1) * 32 + (lane >> 3);
    // input-gradient tile product: rows d_r0 + 32 i (i < 4), columns 4 d_kq .. +3
    const int d_kq = threadIdx.x & 7, d_r0 = threadIdx.x >> 3;
    // weight-gradient blocks
    const int wb = threadIdx.x & 127, whalf = threadIdx.x >> 7;
    const int w_oq = wb & 15, w_kq = wb >> 4;
    float wacc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) wacc[i][j] = 0.f;
    float acc_small = 0.f;       // db2[c] (t<64) | dW1[k][f] (64<=t<160) | db1[k] (160<=t<192)

    for (int sc = blockIdx.x; sc < n_scenes; sc += gridDim.x) {
        const int a = scene_off[sc], n = scene_off[sc + 1] - a;
        if (n <= 1) continue;
        const bool staged = n <= NMAX;
        const size_t poff = (size_t)pair_off[sc];
        __syncthreads();
        if (staged) {
            for (int i = threadIdx.x; i < n * 4; i += MGGAN_THREADS) sX[i] = __ldg(x4 + (size_t)a * 4 + i);
            for (int i = threadIdx.x; i < n * LDU; i += MGGAN_THREADS) sU[i] = __ldg(Us + (size_t)a * LDU + i);
            for (int i = threadIdx.x; i < n * HD; i += MGGAN_THREADS) {
                sHh[(i / HD) * LDHS + (i % HD)] = __ldg(h + (size_t)a * HD + i);
                sdS[(i / HD) * LDHS + (i % HD)] = __ldg(dS + (size_t)a * HD + i);
            }
        }
        __syncthreads();
        const float* pX = staged ? sX : x4 + (size_t)a * 4;
        const float* pU = staged ? sU : Us + (size_t)a * LDU;
        const float* pH = staged ? sHh : h + (size_t)a * HD;
        const float* pdS = staged ? sdS : dS + (size_t)a * HD;
        const int ldh = staged ? LDHS : HD;
        // ---- phase 1a: d sigma_ij = att_ij (dS_i . h_j - sum_j' att_ij' dS_i . h_j'), zero on the diagonal
        for (int il = warp; il < n; il += MGGAN_THREADS / 32) {
            const float* arow = att + poff + (size_t)il * n;
            float* drow = dsig + poff + (size_t)il * n;
            const float* dsi = pdS + (size_t)il * ldh;
            float r = 0.f;
            for (int jl = lane; jl < n; jl += 32) {
                const float* hj = pH + (size_t)jl * ldh;
                float d = 0.f;
#pragma unroll 8
                for (int k = 0; k < HD; ++k) d = fmaf(dsi[k], hj[k], d);
                drow[jl] = d;
                r = fmaf(arow[jl], d, r);
            }
            r = warp_sum(r);
            for (int jl = lane; jl < n; jl += 32) drow[jl] = jl == il ? 0.f : arow[jl] * (drow[jl] - r);
        }
        // ---- phase 1b: dh_j[k] = sum_i att_ij dS_i[k]   (thread = (j, k))
        for (int o = threadIdx.x; o < n * HD; o += MGGAN_THREADS) {
            const int jl = o / HD, k = o - jl * HD;
            const float* ac = att + poff + jl;
            float acc = 0.f;
            for (int il = 0; il < n; ++il) acc = fmaf(__ldg(ac + (size_t)il * n), pdS[(size_t)il * ldh + k], acc);
            dh[(size_t)(a + jl) * HD + k] = acc;
        }
        __syncthreads();                     // dsig complete (written by this CTA only); phase-1 buffers are free
        // ---- phase 2: pair tiles
        const int npairs = n * n;
        for (int p0 = 0; p0 < npairs; p0 += PT) {
            {   // features + layer 1: thread = (pair, half of the 32 units)
                const int pl = threadIdx.x >> 1, half = threadIdx.x & 1;
                const int p = p0 + pl;
                float f1 = 0.f, f2 = 0.f, f3 = 0.f, ds = 0.f;
                int jl = -1;
                if (p < npairs) {
                    jl = p / n;
                    const int il = p - jl * n;
                    ds = dsig[poff + (size_t)il * n + jl];
                    pair_features(ld4(pX + 4 * il), ld4(pX + 4 * jl), f1, f2, f3);
                }
                float* arow = sA1 + pl * LDA1 + half * (F1 / 2);
#pragma unroll
                for (int k = 0; k < F1 / 2; k += 4) {
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 w = ld4(sW1 + 4 * (half * (F1 / 2) + k + q));
                        v[q] = jl >= 0 ? fmaxf(fmaf(w.x, f1, fmaf(w.y, f2, fmaf(w.z, f3, w.w))), 0.f) : 0.f;
                    }
                    st4(arow + k, make_float4(v[0], v[1], v[2], v[3]));
                }
                if (half == 0) {
                    st4(sF + pl * 4, make_float4(f1, f2, f3, 0.f));
                    sDs[pl] = ds;
                    sJ[pl] = jl;
                }
            }
            __syncthreads();
            {   // S = A1 W2^T + b2 -> D2, dU
                float acc[8][4];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const float bv = sb2[u + 16 * jj];
#pragma unroll
                    for (int i = 0; i < 8; ++i) acc[i][jj] = bv;
                }
                tile_rowdot<8, 4, F1>(acc, sA1, LDA1, rl, 4, sW2, LDW2, u, 16);
                int curj = -1;
                float uj[4] = {0.f, 0.f, 0.f, 0.f}, run[4] = {0.f, 0.f, 0.f, 0.f}, run_s = 0.f;
                const bool sown = u == 0;             // one column owner per row also carries the s_j term
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int r = rl + 4 * i;
                    const int jl = sJ[r];
                    const float ds = sDs[r];
                    if (jl != curj) {
                        if (curj >= 0) {
                            float* du = dUs + (size_t)(a + curj) * LDU;
#pragma unroll
                            for (int jj = 0; jj < 4; ++jj)
                                if (run[jj] != 0.f) atomicAdd(du + u + 16 * jj, run[jj]);
                            if (sown && run_s != 0.f) atomicAdd(du + F2, run_s);
                        }
                        curj = jl;
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj) { run[jj] = 0.f; uj[jj] = jl >= 0 ? pU[(size_t)jl * LDU + u + 16 * jj] : 0.f; }
                        run_s = 0.f;
                    }
                    float d2[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const float sv = acc[i][jj];
                        run[jj] = fmaf(ds, fmaxf(sv, 0.f), run[jj]);
                        d2[jj] = sv > 0.f ? ds * uj[jj] : 0.f;
                        sD2[r * LDA2 + u + 16 * jj] = d2[jj];
                    }
                    run_s += ds;
                }
                if (curj >= 0) {
                    float* du = dUs + (size_t)(a + curj) * LDU;
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                        if (run[jj] != 0.f) atomicAdd(du + u + 16 * jj, run[jj]);
                    if (sown && run_s != 0.f) atomicAdd(du + F2, run_s);
                }
            }
            __syncthreads();
            {   // dA1 = (D2 W2) [A1 > 0]
                float acc[4][4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { acc[i][0] = 0.f; acc[i][1] = 0.f; acc[i][2] = 0.f; acc[i][3] = 0.f; }
                tile_dgrad<4, F2>(acc, sD2, LDA2, d_r0, 32, sW2, LDW2, d_kq * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = d_r0 + 32 * i;
                    const float4 a1 = ld4(sA1 + r * LDA1 + d_kq * 4);
                    st4(sDA1 + r * LDA1 + d_kq * 4,
                        make_float4(a1.x > 0.f ? acc[i][0] : 0.f, a1.y > 0.f ? acc[i][1] : 0.f,
                                    a1.z > 0.f ? acc[i][2] : 0.f, a1.w > 0.f ? acc[i][3] : 0.f));
                }
            }
            __syncthreads();
            tile_wgrad<PT / 2>(wacc, sD2 + whalf * (PT / 2) * LDA2, LDA2, w_oq * 4, sA1 + whalf * (PT / 2) * LDA1, LDA1,
                               w_kq * 4);
            {
                const int t = threadIdx.x;
                if (t < 64) {
                    for (int r = 0; r < PT; ++r) acc_small += sD2[r * LDA2 + t];
                } else if (t < 160) {
                    int k = (t - 64) / 3, f = (t - 64) % 3;
                    for (int r = 0; r < PT; ++r) acc_small = fmaf(sDA1[r * LDA1 + k], sF[r * 4 + f], acc_small);
                } else if (t < 192) {
                    int k = t - 160;
                    for (int r = 0; r < PT; ++r) acc_small += sDA1[r * LDA1 + k];
                }
            }
            __syncthreads();
        }
    }
    atomic_block44(dW2, F1, w_oq * 4, w_kq * 4, wacc);
    {
        const int t = threadIdx.x;
        if (t < 64) atomicAdd(db2 + t, acc_small);
        else if (t < 160) atomicAdd(dW1 + (t - 64), acc_small);
        else if (t < 192) atomicAdd(db1 + (t - 160), acc_small);
    }
}

template <int HD>
size_t soc_fwd_smem() { return sizeof(float) * (F2 * LDW2 + F1 * 4 + F2 + NMAX * 4 + NMAX * LDU + NMAX * (HD + 1)); }
template <int HD>
size_t soc_bwd_smem() {
    static_assert(2 * NMAX * (HD + 1) <= PT * LDA1 + PT * LDA2, "phase-1 buffers alias the pair tiles");
    return sizeof(float) * (F2 * LDW2 + F1 * 4 + F2 + NMAX * 4 + ((NMAX * LDU + 3) & ~3) + 2 * PT * LDA1 + PT * LDA2 +
                            PT * 4 + 2 * PT);
}

int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int HD>
int launch_fwd(const float* x4, const float* h, const float* Us, const int* so, const int* po, int ns, const float* W1,
               const float* b1, const float* W2, const float* b2, float* S, float* att, cudaStream_t st) {
    size_t sm = soc_fwd_smem<HD>();
    cudaFuncSetAttribute(social_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int grid = ns < sm_count() * 4 ? ns : sm_count() * 4;
    social_fwd_kernel<HD><<<grid, MGGAN_THREADS, sm, st>>>(x4, h, Us, so, po, ns, W1, b1, W2, b2, S, att);
    return mggan_check_launch("social_attn_fwd");
}
template <int HD>
int launch_bwd(const float* x4, const float* h, const float* Us, const int* so, const int* po, int ns, const float* W1,
               const float* b1, const float* W2, const float* b2, const float* att, const float* dS, float* dsig,
               float* dh, float* dUs, float* dW1, float* db1, float* dW2, float* db2, cudaStream_t st) {
    size_t sm = soc_bwd_smem<HD>();
    cudaFuncSetAttribute(social_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int grid = ns < sm_count() * 2 ? ns : sm_count() * 2;
    social_bwd_kernel<HD><<<grid, MGGAN_THREADS, sm, st>>>(x4, h, Us, so, po, ns, W1, b1, W2, b2, att, dS, dsig, dh, dUs,
                                                           dW1, db1, dW2, db2);
    return mggan_check_launch("social_attn_bwd");
}

}  // namespace

extern "C" int mggan_social_attn_fwd(const float* x4, const float* h, int HD, const float* Us, const int* scene_off,
                                     const int* pair_off, int n_scenes, const float* W1, const float* b1,
                                     const float* W2, const float* b2, float* S, float* att, cudaStream_t stream) {
    MGGAN_REQUIRE(HD == 32 || HD == 64, "mggan_social_attn_fwd: hidden size %d not built (32 or 64)", HD);
    if (n_scenes <= 0) return MGGAN_OK;
    return HD == 32 ? launch_fwd<32>(x4, h, Us, scene_off, pair_off, n_scenes, W1, b1, W2, b2, S, att, stream)
                    : launch_fwd<64>(x4, h, Us, scene_off, pair_off, n_scenes, W1, b1, W2, b2, S, att, stream);
}

extern "C" int mggan_social_attn_bwd(const float* x4, const float* h, int HD, const float* Us, const int* scene_off,
                                     const int* pair_off, int n_scenes, const float* W1, const float* b1,
                                     const float* W2, const float* b2, const float* att, const float* dS, float* dsig,
                                     float* dh, float* dUs, float* dW1, float* db1, float* dW2, float* db2,
                                     cudaStream_t stream) {
    MGGAN_REQUIRE(HD == 32 || HD == 64, "mggan_social_attn_bwd: hidden size %d not built (32 or 64)", HD);
    if (n_scenes <= 0) return MGGAN_OK;
    return HD == 32 ? launch_bwd<32>(x4, h, Us, scene_off, pair_off, n_scenes, W1, b1, W2, b2, att, dS, dsig, dh, dUs,
                                     dW1, db1, dW2, db2, stream)
                    : launch_bwd<64>(x4, h, Us, scene_off, pair_off, n_scenes, W1, b1, W2, b2, att, dS, dsig, dh, dUs,
                                     dW1, db1, dW2, db2, stream);
}
