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

// Forward.  Per scene: (1) sigma for the n^2 ordered pairs in tiles of PT pairs (pair p = i n + j, so a tile is a
// contiguous piece of the scene's attention matrix): features + layer 1 thread-per-(pair, half), layer 2 as a warp-level
// 3 x TF32 tensor-pipe product (warp = 16 pairs x 64 units, W2 pre-split), sigma = relu(S) . u_j + s_j reduced over the four
// lanes that share a row; (2) per agent (warp): softmax over its row and S_i = sum_j att_ij h_j.
// (The lane-per-pair FP32 form spent 2,048 FMAs and 512 LDS.128 per pair-lane on layer 2: 0.12 ms per launch.)
template <int HD>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
social_fwd_kernel(const float* __restrict__ x4, const float* __restrict__ h, const float* __restrict__ Us,
                  const int* __restrict__ scene_off, const int* __restrict__ pair_off, int n_scenes,
                  const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                  const float* __restrict__ b2, float* __restrict__ S, float* __restrict__ att) {
    constexpr int LDHS = HD + 1;
    extern __shared__ __align__(16) float smem[];
    float* sW2 = smem;                       // [F2][LDW2]  TF32 hi plane of W2
    float* sW2l = sW2 + F2 * LDW2;           // [F2][LDW2]  lo plane
    float* sW1 = sW2l + F2 * LDW2;           // [F1][4]
    float* sb2 = sW1 + F1 * 4;               // [F2]
    float* sX = sb2 + F2;                    // [NMAX][4]
    float* sU = sX + NMAX * 4;               // [NMAX][LDU]
    float* sHh = sU + ((NMAX * LDU + 3) & ~3);   // [NMAX][LDHS]
    float* sA1 = sHh + ((NMAX * LDHS + 3) & ~3); // [PT][LDA1]
    stage_pair_weights(sW1, sW2, sb2, W1, b1, W2, b2);
    __syncthreads();
    for (int i = threadIdx.x; i < F2 * LDW2; i += MGGAN_THREADS) {
        uint32_t hi, lo;
        tf32_split(sW2[i], hi, lo);
        sW2[i] = __uint_as_float(hi);
        sW2l[i] = __uint_as_float(lo);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g8 = lane >> 2, t4 = lane & 3;
    const int m0 = warp * 16;

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
        // ---- (1) sigma of every ordered pair, PT pairs at a time; every stage of a tile is local to the warp's 16 pairs
        const int npairs = n * n;
        for (int p0 = 0; p0 < npairs; p0 += PT) {
            {   // features + layer 1: thread = (pair, half of the 32 units)
                const int pl = threadIdx.x >> 1, half = threadIdx.x & 1;
                const int p = p0 + pl;
                float f1 = 0.f, f2 = 0.f, f3 = 0.f;
                if (p < npairs) {
                    const int il = p / n, jl = p - il * n;
                    pair_features(ld4(pX + 4 * il), ld4(pX + 4 * jl), f1, f2, f3);
                }
                float* arow = sA1 + pl * LDA1 + half * (F1 / 2);
#pragma unroll
                for (int k = 0; k < F1 / 2; k += 4) {
                    float v[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float4 w = ld4(sW1 + 4 * (half * (F1 / 2) + k + q));
                        v[q] = fmaxf(fmaf(w.x, f1, fmaf(w.y, f2, fmaf(w.z, f3, w.w))), 0.f);
                    }
                    st4(arow + k, make_float4(v[0], v[1], v[2], v[3]));
                }
            }
            __syncwarp();
            {   // S = A1 W2^T + b2 (warp = pairs m0 .. m0 + 15), then sigma = relu(S) . u_j + s_j
                float acc[8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float ba = sb2[8 * j + 2 * t4], bb = sb2[8 * j + 2 * t4 + 1];
                    acc[j][0] = ba; acc[j][1] = bb; acc[j][2] = ba; acc[j][3] = bb;
                }
#pragma unroll
                for (int k0 = 0; k0 < F1; k0 += 8) {
                    const float* pa = sA1 + (m0 + g8) * LDA1 + k0 + t4;
                    uint32_t ah[4], al[4];
                    tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8 * LDA1], ah[1], al[1]);
                    tf32_split(pa[4], ah[2], al[2]); tf32_split(pa[8 * LDA1 + 4], ah[3], al[3]);
                    const int ob = g8 * LDW2 + k0 + t4;
#pragma unroll
                    for (int j0 = 0; j0 < 8; j0 += 4) {
                        uint32_t bh[4][2], bl[4][2];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int o = ob + 8 * (j0 + j) * LDW2;
                            bh[j][0] = __float_as_uint(sW2[o]); bh[j][1] = __float_as_uint(sW2[o + 4]);
                            bl[j][0] = __float_as_uint(sW2l[o]); bl[j][1] = __float_as_uint(sW2l[o + 4]);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j0 + j], ah, bh[j][0], bh[j][1]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j0 + j], al, bh[j][0], bh[j][1]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j0 + j], ah, bl[j][0], bl[j][1]);
                    }
                }
                // rows m0 + g (c0, c1) and m0 + g + 8 (c2, c3), columns 8 j + 2 t + {0, 1}
                const int pa_ = p0 + m0 + g8, pb_ = pa_ + 8;
                const int ia = pa_ < npairs ? pa_ / n : 0, ja = pa_ < npairs ? pa_ - ia * n : 0;
                const int ib = pb_ < npairs ? pb_ / n : 0, jb = pb_ < npairs ? pb_ - ib * n : 0;
                const float* ua = pU + (size_t)ja * LDU;
                const float* ub = pU + (size_t)jb * LDU;
                float sa = 0.f, sb = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = 8 * j + 2 * t4;
                    sa = fmaf(fmaxf(acc[j][0], 0.f), ua[c], fmaf(fmaxf(acc[j][1], 0.f), ua[c + 1], sa));
                    sb = fmaf(fmaxf(acc[j][2], 0.f), ub[c], fmaf(fmaxf(acc[j][3], 0.f), ub[c + 1], sb));
                }
                sa += __shfl_xor_sync(0xffffffffu, sa, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 1);
                sa += __shfl_xor_sync(0xffffffffu, sa, 2); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
                if (t4 == 0) {
                    if (pa_ < npairs) att[poff + pa_] = ia == ja ? -1000.f : sa + ua[F2];
                    if (pb_ < npairs) att[poff + pb_] = ib == jb ? -1000.f : sb + ub[F2];
                }
            }
            __syncwarp();                    // the warp's rows of sA1 are rewritten by the next tile's layer 1
        }
        __syncthreads();                     // sigma of the whole scene is in `att` (written by this CTA only)
        // ---- (2) softmax over each agent's row and the weighted sum of the neighbours' hidden states
        for (int il = warp; il < n; il += MGGAN_THREADS / 32) {
            float* arow = att + poff + (size_t)il * n;
            float mx = -INFINITY;
            for (int jl = lane; jl < n; jl += 32) mx = fmaxf(mx, arow[jl]);
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
// (pair p -> j = p / n, i = p % n); the three dense pieces are warp-level 3 x TF32 tensor-pipe products (common.cuh;
// as FP32 register tiles they were 85 % of the kernel's samples at the FMA-issue floor):
//     S  = A1 W2^T + b2        (PT x 64, K = 32)     recomputed pre-activation of layer 2; warp = 16 pairs x 64
//     D2 = [S > 0] ds u_j ,  dU_j += ds relu(S)      (16-row tiles of one neighbour j: shuffle-reduced, one atomic per column)
//     dA1 = D2 W2 [A1 > 0]     (PT x 32, K = 64)     warp = 16 pairs x 32
//     dW2 += D2^T A1           (64 x 32, K = PT)     warp = 16 outputs x 16 inputs, C fragments kept across tiles
//     db2 += colsum D2, dW1 += dA1^T F, db1 += colsum dA1   (FP32, kept in registers across tiles)
// W2 is staged pre-split as (hi, lo) planes; the activations are split where they are read.
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
    float* sW2 = smem;                       // [F2][LDW2]  TF32 hi plane of W2
    float* sW2l = sW2 + F2 * LDW2;           // [F2][LDW2]  lo plane (W2 - hi)
    float* sW1 = sW2l + F2 * LDW2;           // [F1][4]
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
    __syncthreads();
    for (int i = threadIdx.x; i < F2 * LDW2; i += MGGAN_THREADS) {       // split W2 in place: (hi, lo) planes
        uint32_t hi, lo;
        tf32_split(sW2[i], hi, lo);
        sW2[i] = __uint_as_float(hi);
        sW2l[i] = __uint_as_float(lo);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g8 = lane >> 2, t4 = lane & 3;                // MMA fragment coordinates
    const int m0 = warp * 16;                               // the warp's 16 pairs of a tile
    const int wm0 = (warp >> 1) * 16, wn0 = (warp & 1) * 16;    // dW2 block of the warp: outputs wm0.., inputs wn0..
    float wacc[2][4];                                       // dW2 C fragments (n-tiles wn0, wn0 + 8), kept across tiles
#pragma unroll
    for (int i = 0; i < 2; ++i)
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
        // the scene's attention matrix (phase 1b reads it by columns: n dependent strided global loads per thread were
        // 9 % of the kernel's samples) -> shared memory, row stride n + 1; sDA1 is free until phase 2
        const float* pAtt = att + poff;
        int lda_att = n;
        if (staged) {
            for (int i = threadIdx.x; i < n * n; i += MGGAN_THREADS) sDA1[(i / n) * (n + 1) + i % n] = __ldg(att + poff + i);
            pAtt = sDA1;
            lda_att = n + 1;
            __syncthreads();
        }
        // ---- phase 1a: d sigma_ij = att_ij (dS_i . h_j - sum_j' att_ij' dS_i . h_j'), zero on the diagonal
        for (int il = warp; il < n; il += MGGAN_THREADS / 32) {
            const float* arow = pAtt + (size_t)il * lda_att;
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
            const float* ac = pAtt + jl;
            float acc = 0.f;
#pragma unroll 4
            for (int il = 0; il < n; ++il) acc = fmaf(ac[(size_t)il * lda_att], pdS[(size_t)il * ldh + k], acc);
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
            __syncwarp();                        // warp w produced exactly the 16 pairs (rows m0 ..) it consumes next
            {   // S = A1 W2^T + b2 -> D2, dU.  warp = pairs m0 .. m0 + 15; A (row g, k t) = sA1[(m0 + g) LDA1 + k0 + t],
                // B (k t, n g) = W2[n0 + g][k0 + t]: both on banks 4 g + t
                float acc[8][4];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float ba = sb2[8 * j + 2 * t4], bb = sb2[8 * j + 2 * t4 + 1];
                    acc[j][0] = ba; acc[j][1] = bb; acc[j][2] = ba; acc[j][3] = bb;
                }
#pragma unroll
                for (int k0 = 0; k0 < F1; k0 += 8) {
                    const float* pa = sA1 + (m0 + g8) * LDA1 + k0 + t4;
                    uint32_t ah[4], al[4];
                    tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8 * LDA1], ah[1], al[1]);
                    tf32_split(pa[4], ah[2], al[2]); tf32_split(pa[8 * LDA1 + 4], ah[3], al[3]);
                    const int ob = g8 * LDW2 + k0 + t4;
#pragma unroll
                    for (int j0 = 0; j0 < 8; j0 += 4) {      // MMAs product-major over four accumulators at a time
                        uint32_t bh[4][2], bl[4][2];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int o = ob + 8 * (j0 + j) * LDW2;
                            bh[j][0] = __float_as_uint(sW2[o]); bh[j][1] = __float_as_uint(sW2[o + 4]);
                            bl[j][0] = __float_as_uint(sW2l[o]); bl[j][1] = __float_as_uint(sW2l[o + 4]);
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j0 + j], ah, bh[j][0], bh[j][1]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j0 + j], al, bh[j][0], bh[j][1]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j0 + j], ah, bl[j][0], bl[j][1]);
                    }
                }
                // the thread holds rows m0 + g, m0 + g + 8 and columns 8 j + 2 t + {0, 1}
                const int ra = m0 + g8, rb = ra + 8;
                const int ja = sJ[ra], jb = sJ[rb];
                const float dsa = sDs[ra], dsb = sDs[rb];
                const bool one_j = sJ[m0] == sJ[m0 + 15] && sJ[m0] >= 0;      // the tile's 16 pairs share their neighbour (warp-uniform)
                const float* ua = pU + (size_t)(ja >= 0 ? ja : 0) * LDU;
                const float* ub = pU + (size_t)(jb >= 0 ? jb : 0) * LDU;
                float* dua = dUs + (size_t)(a + (ja >= 0 ? ja : 0)) * LDU;
                float* dub = dUs + (size_t)(a + (jb >= 0 ? jb : 0)) * LDU;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = 8 * j + 2 * t4;
                    const float s0 = acc[j][0], s1 = acc[j][1], s2 = acc[j][2], s3 = acc[j][3];
                    *reinterpret_cast<float2*>(sD2 + ra * LDA2 + c) =
                        make_float2(s0 > 0.f && ja >= 0 ? dsa * ua[c] : 0.f, s1 > 0.f && ja >= 0 ? dsa * ua[c + 1] : 0.f);
                    *reinterpret_cast<float2*>(sD2 + rb * LDA2 + c) =
                        make_float2(s2 > 0.f && jb >= 0 ? dsb * ub[c] : 0.f, s3 > 0.f && jb >= 0 ? dsb * ub[c + 1] : 0.f);
                    // dU_j[c] += ds relu(S)
                    float v0 = dsa * fmaxf(s0, 0.f), v1 = dsa * fmaxf(s1, 0.f), v2 = dsb * fmaxf(s2, 0.f), v3 = dsb * fmaxf(s3, 0.f);
                    if (one_j) {
                        v0 += v2; v1 += v3;
                        v0 += __shfl_xor_sync(0xffffffffu, v0, 4); v1 += __shfl_xor_sync(0xffffffffu, v1, 4);
                        v0 += __shfl_xor_sync(0xffffffffu, v0, 8); v1 += __shfl_xor_sync(0xffffffffu, v1, 8);
                        v0 += __shfl_xor_sync(0xffffffffu, v0, 16); v1 += __shfl_xor_sync(0xffffffffu, v1, 16);
                        if (g8 == 0) { atomicAdd(dua + c, v0); atomicAdd(dua + c + 1, v1); }
                    } else {
                        if (ja >= 0) { if (v0 != 0.f) atomicAdd(dua + c, v0); if (v1 != 0.f) atomicAdd(dua + c + 1, v1); }
                        if (jb >= 0) { if (v2 != 0.f) atomicAdd(dub + c, v2); if (v3 != 0.f) atomicAdd(dub + c + 1, v3); }
                    }
                }
                {   // the s_j term: dU_j[F2] += ds
                    float va = ja >= 0 ? dsa : 0.f, vb = jb >= 0 ? dsb : 0.f;
                    if (one_j) {
                        va += vb;
                        va += __shfl_xor_sync(0xffffffffu, va, 4);
                        va += __shfl_xor_sync(0xffffffffu, va, 8);
                        va += __shfl_xor_sync(0xffffffffu, va, 16);
                        if (lane == 0) atomicAdd(dua + F2, va);
                    } else if (t4 == 0) {
                        if (va != 0.f) atomicAdd(dua + F2, va);
                        if (vb != 0.f) atomicAdd(dub + F2, vb);
                    }
                }
            }
            __syncwarp();                        // D2 rows m0 .. m0 + 15 were written by this warp
            {   // dA1 = (D2 W2) [A1 > 0]: warp = pairs m0 .. m0 + 15 x 32 inputs, K = 64 outputs;
                // B (k t, n g) = W2[k0 + t][n0 + g] (banks 4 t + g: 2-way conflicts, the price of one copy of W2 for both products)
                float acc[4][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { acc[j][0] = 0.f; acc[j][1] = 0.f; acc[j][2] = 0.f; acc[j][3] = 0.f; }
#pragma unroll
                for (int k0 = 0; k0 < F2; k0 += 8) {
                    const float* pa = sD2 + (m0 + g8) * LDA2 + k0 + t4;
                    uint32_t ah[4], al[4];
                    tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8 * LDA2], ah[1], al[1]);
                    tf32_split(pa[4], ah[2], al[2]); tf32_split(pa[8 * LDA2 + 4], ah[3], al[3]);
                    const int ob = (k0 + t4) * LDW2 + g8;
                    uint32_t bh[4][2], bl[4][2];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        bh[j][0] = __float_as_uint(sW2[ob + 8 * j]); bh[j][1] = __float_as_uint(sW2[ob + 4 * LDW2 + 8 * j]);
                        bl[j][0] = __float_as_uint(sW2l[ob + 8 * j]); bl[j][1] = __float_as_uint(sW2l[ob + 4 * LDW2 + 8 * j]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j], ah, bh[j][0], bh[j][1]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j], al, bh[j][0], bh[j][1]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) mma_tf32_16x8x8(acc[j], ah, bl[j][0], bl[j][1]);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int c = 8 * j + 2 * t4;
                    const float2 a1a = *reinterpret_cast<const float2*>(sA1 + (m0 + g8) * LDA1 + c);
                    const float2 a1b = *reinterpret_cast<const float2*>(sA1 + (m0 + g8 + 8) * LDA1 + c);
                    *reinterpret_cast<float2*>(sDA1 + (m0 + g8) * LDA1 + c) =
                        make_float2(a1a.x > 0.f ? acc[j][0] : 0.f, a1a.y > 0.f ? acc[j][1] : 0.f);
                    *reinterpret_cast<float2*>(sDA1 + (m0 + g8 + 8) * LDA1 + c) =
                        make_float2(a1b.x > 0.f ? acc[j][2] : 0.f, a1b.y > 0.f ? acc[j][3] : 0.f);
                }
            }
            __syncthreads();
            {   // dW2 += D2^T A1: warp = outputs wm0 .. wm0 + 15 x inputs wn0 .. wn0 + 15, K = the tile's pairs;
                // A (row g, k t) = sD2[(k0 + t) LDA2 + wm0 + g], B (k t, n g) = sA1[(k0 + t) LDA1 + n + g]
#pragma unroll 4
                for (int k0 = 0; k0 < PT; k0 += 8) {
                    const float* pa = sD2 + (k0 + t4) * LDA2 + wm0 + g8;
                    const float* pb = sA1 + (k0 + t4) * LDA1 + wn0 + g8;
                    uint32_t ah[4], al[4];
                    tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8], ah[1], al[1]);
                    tf32_split(pa[4 * LDA2], ah[2], al[2]); tf32_split(pa[4 * LDA2 + 8], ah[3], al[3]);
                    uint32_t bh[4], bl[4];
                    tf32_split(pb[0], bh[0], bl[0]); tf32_split(pb[4 * LDA1], bh[1], bl[1]);
                    tf32_split(pb[8], bh[2], bl[2]); tf32_split(pb[4 * LDA1 + 8], bh[3], bl[3]);
                    mma_tf32_16x8x8(wacc[0], ah, bh[0], bh[1]); mma_tf32_16x8x8(wacc[1], ah, bh[2], bh[3]);
                    mma_tf32_16x8x8(wacc[0], al, bh[0], bh[1]); mma_tf32_16x8x8(wacc[1], al, bh[2], bh[3]);
                    mma_tf32_16x8x8(wacc[0], ah, bl[0], bl[1]); mma_tf32_16x8x8(wacc[1], ah, bl[2], bl[3]);
                }
            }
            {
                const int t = threadIdx.x;
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;      // four partial sums: a 128-long dependent chain otherwise
                if (t < 64) {
#pragma unroll 4
                    for (int r = 0; r < PT; r += 4) {
                        p0 += sD2[r * LDA2 + t]; p1 += sD2[(r + 1) * LDA2 + t]; p2 += sD2[(r + 2) * LDA2 + t]; p3 += sD2[(r + 3) * LDA2 + t];
                    }
                } else if (t < 160) {
                    int k = (t - 64) / 3, f = (t - 64) % 3;
#pragma unroll 4
                    for (int r = 0; r < PT; r += 4) {
                        p0 = fmaf(sDA1[r * LDA1 + k], sF[r * 4 + f], p0); p1 = fmaf(sDA1[(r + 1) * LDA1 + k], sF[(r + 1) * 4 + f], p1);
                        p2 = fmaf(sDA1[(r + 2) * LDA1 + k], sF[(r + 2) * 4 + f], p2); p3 = fmaf(sDA1[(r + 3) * LDA1 + k], sF[(r + 3) * 4 + f], p3);
                    }
                } else if (t < 192) {
                    int k = t - 160;
#pragma unroll 4
                    for (int r = 0; r < PT; r += 4) {
                        p0 += sDA1[r * LDA1 + k]; p1 += sDA1[(r + 1) * LDA1 + k]; p2 += sDA1[(r + 2) * LDA1 + k]; p3 += sDA1[(r + 3) * LDA1 + k];
                    }
                }
                acc_small += (p0 + p1) + (p2 + p3);
            }
            __syncthreads();
        }
    }
    // dW2 C fragments: c0, c1 -> output wm0 + g, inputs n + 2t, n + 2t + 1; c2, c3 -> output wm0 + g + 8
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        float* dst = dW2 + (size_t)(wm0 + g8) * F1 + wn0 + 8 * i + 2 * t4;
        atomicAdd(dst, wacc[i][0]); atomicAdd(dst + 1, wacc[i][1]);
        atomicAdd(dst + 8 * F1, wacc[i][2]); atomicAdd(dst + 8 * F1 + 1, wacc[i][3]);
    }
    {
        const int t = threadIdx.x;
        if (t < 64) atomicAdd(db2 + t, acc_small);
        else if (t < 160) atomicAdd(dW1 + (t - 64), acc_small);
        else if (t < 192) atomicAdd(db1 + (t - 160), acc_small);
    }
}

template <int HD>
size_t soc_fwd_smem() { return sizeof(float) * (2 * F2 * LDW2 + F1 * 4 + F2 + NMAX * 4 + ((NMAX * LDU + 3) & ~3) + ((NMAX * (HD + 1) + 3) & ~3) + PT * LDA1); }
template <int HD>
size_t soc_bwd_smem() {
    static_assert(2 * NMAX * (HD + 1) <= PT * LDA1 + PT * LDA2, "phase-1 buffers alias the pair tiles");
    return sizeof(float) * (2 * F2 * LDW2 + F1 * 4 + F2 + NMAX * 4 + ((NMAX * LDU + 3) & ~3) + 2 * PT * LDA1 + PT * LDA2 +
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
    int grid = ns < sm_count() * 2 ? ns : sm_count() * 2;
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
