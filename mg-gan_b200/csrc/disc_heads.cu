// Discriminator heads over k samples per agent with the per-agent part of the first layer hoisted
// (reference: MultiDiscriminatorTrajectory.forward mggan/model/modules/discriminators.py:178-219:
// classifier_inp = cat[soc, enc, scene] -> discs[0] (Linear, LeakyReLU(0.2), Linear, Sigmoid, eps squash :203-204)
// and gen_id_reconstructor (Linear, LeakyReLU(0.2), Linear) :103-108,209-217).
//
// The classifier input of row (sample s, agent i) is [soc | in_enc | pred_enc | scene].  Only pred_enc (32 of the
// 192 columns) depends on the sample, and soc is non-zero for s == 0 only (the reference's seq_start_end * k quirk,
// SURVEY.md 3.3).  With the first layers of both heads stacked (NZ = 2 HH outputs, or HH without the classifier):
//     z[s, i] = base[i] + (s == 0) * soc0[i] + W1p pe[s, i]
// base = W1[:, in_enc | scene cols] [in_enc_i | scene_i] + b1 and soc0 = W1[:, soc cols] soc_i are N-row products the
// caller computes once per agent; this kernel does the per-sample K = 32 product, both second layers and the
// sigmoid, 64 rows per CTA tile, weights resident in shared memory.  The backward produces input gradients only
// (d pe, d soc0, optionally d base): it serves the generator step, where the discriminator is frozen; the
// discriminator step (k = 1, N rows) trains these weights through the dense-layer kernels.
#include "common.cuh"

namespace {

constexpr int KP = 32;            // pred_enc width (h_dim // 2)
constexpr int LDK = KP + 4;       // 36
constexpr int ROWS = 64;
constexpr int GMAX = 16;
constexpr float SLOPE = 0.2f;
constexpr float D_EPS = 1e-7f;    // discriminators.py:110

template <int TOD, bool HASG>
struct Cfg {
    static constexpr int HH = TOD * 32;                   // hidden width of one head (96 or 64)
    static constexpr int TO = HASG ? 2 * TOD : TOD;       // 32-column groups of the stacked first layer
    static constexpr int NZ = TO * 32;
    static constexpr int LDA = NZ + 4;
    static constexpr int LDW2 = HH + 4;
};

template <int TOD, bool HASG>
__device__ __forceinline__ void stage_weights(float* sW1, float* sW2, const float* __restrict__ W1p,
                                              const float* __restrict__ Wd2, const float* __restrict__ Wg2, int G) {
    using C = Cfg<TOD, HASG>;
    stage_matrix(sW1, LDK, W1p, C::NZ, KP);
    for (int i = threadIdx.x; i < C::HH; i += MGGAN_THREADS) sW2[i] = __ldg(Wd2 + i);
    if (HASG)
        for (int i = threadIdx.x; i < G * C::HH; i += MGGAN_THREADS) sW2[(1 + i / C::HH) * C::LDW2 + i % C::HH] = __ldg(Wg2 + i);
}

// sA[r][:] = base[i] + (s == 0) soc0[i]; sPe[r][:] = pe[row]; rows past the end are zero.
template <int TOD, bool HASG>
__device__ __forceinline__ void stage_tile(float* sA, float* sPe, const float* __restrict__ pe,
                                           const float* __restrict__ base, const float* __restrict__ soc0, long long row0,
                                           long long R, int n) {
    using C = Cfg<TOD, HASG>;
    constexpr int Q = C::NZ / 4;
    for (int idx = threadIdx.x; idx < ROWS * Q; idx += MGGAN_THREADS) {
        int r = idx / Q, c = (idx - r * Q) * 4;
        long long row = row0 + r;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < R) {
            int i = (int)(row % n);
            v = __ldg(reinterpret_cast<const float4*>(base + (size_t)i * C::NZ + c));
            if (row < n) {
                float4 s = __ldg(reinterpret_cast<const float4*>(soc0 + (size_t)i * C::NZ + c));
                v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
            }
        }
        st4(sA + r * C::LDA + c, v);
    }
    for (int idx = threadIdx.x; idx < ROWS * (KP / 4); idx += MGGAN_THREADS) {
        int r = idx / (KP / 4), c = (idx - r * (KP / 4)) * 4;
        long long row = row0 + r;
        float4 v = row < R ? __ldg(reinterpret_cast<const float4*>(pe + (size_t)row * KP + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        st4(sPe + r * LDK + c, v);
    }
}

// Backward tiling: a tile holds ALL k samples of ROWS / k consecutive agents (tile row r = a k + s -> global row s n + agent),
// so that d base[agent] = sum_s dz[s, agent] is a sum over rows of the same tile: round 1 added every (row, column) of dz to
// d_base with its own atomicAdd (63 M atomics per generator step at k = 20, n = 16,384).
struct RowMap {
    int k, n, apt;                 // apt = agents per tile
    long long agent0;              // first agent of the tile
    __device__ __forceinline__ bool map(int r, long long& row, int& ag, int& smp) const {
        const int a = r / k;
        smp = r - a * k;
        ag = (int)(agent0 + a);
        row = (long long)smp * n + ag;
        return a < apt && ag < n;
    }
};

template <int TOD, bool HASG>
__device__ __forceinline__ void stage_tile_grouped(float* sA, float* sPe, const float* __restrict__ pe,
                                                   const float* __restrict__ base, const float* __restrict__ soc0,
                                                   const RowMap& rm) {
    using C = Cfg<TOD, HASG>;
    constexpr int Q = C::NZ / 4;
    for (int idx = threadIdx.x; idx < ROWS * Q; idx += MGGAN_THREADS) {
        int r = idx / Q, c = (idx - r * Q) * 4;
        long long row; int ag, smp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rm.map(r, row, ag, smp)) {
            v = __ldg(reinterpret_cast<const float4*>(base + (size_t)ag * C::NZ + c));
            if (smp == 0) {
                float4 q = __ldg(reinterpret_cast<const float4*>(soc0 + (size_t)ag * C::NZ + c));
                v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
            }
        }
        st4(sA + r * C::LDA + c, v);
    }
    for (int idx = threadIdx.x; idx < ROWS * (KP / 4); idx += MGGAN_THREADS) {
        int r = idx / (KP / 4), c = (idx - r * (KP / 4)) * 4;
        long long row; int ag, smp;
        float4 v = rm.map(r, row, ag, smp) ? __ldg(reinterpret_cast<const float4*>(pe + (size_t)row * KP + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        st4(sPe + r * LDK + c, v);
    }
}

template <int TOD, bool HASG>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
disc_heads_fwd_kernel(const float* __restrict__ pe, int n, int k, const float* __restrict__ base,
                      const float* __restrict__ soc0, const float* __restrict__ W1p, const float* __restrict__ Wd2,
                      const float* __restrict__ bd2, const float* __restrict__ Wg2, const float* __restrict__ bg2, int G,
                      float* __restrict__ p, float* __restrict__ branch) {
    using C = Cfg<TOD, HASG>;
    extern __shared__ __align__(16) float smem[];
    float* sW1 = smem;                          // [NZ][LDK]
    float* sPe = sW1 + C::NZ * LDK;             // [ROWS][LDK]
    float* sA = sPe + ROWS * LDK;               // [ROWS][LDA]
    float* sW2 = sA + ROWS * C::LDA;            // [1 + G][LDW2]
    stage_weights<TOD, HASG>(sW1, sW2, W1p, Wd2, Wg2, G);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = (warp & 3) * 8 + (lane & 7);
    const int rl = (warp >> 2) * 32 + (lane >> 3);
    const int prow = threadIdx.x >> 2, q = threadIdx.x & 3;
    const long long R = (long long)n * k;
    const long long n_tiles = (R + ROWS - 1) / ROWS;
    const float b_d = __ldg(bd2);

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const long long row0 = tile * ROWS;
        __syncthreads();
        stage_tile<TOD, HASG>(sA, sPe, pe, base, soc0, row0, R, n);
        __syncthreads();
        {
            float acc[8][C::TO];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < C::TO; ++j) acc[i][j] = sA[(rl + 4 * i) * C::LDA + u + 32 * j];
            tile_rowdot<8, C::TO, KP>(acc, sPe, LDK, rl, 4, sW1, LDK, u, 32);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < C::TO; ++j) sA[(rl + 4 * i) * C::LDA + u + 32 * j] = lrelu_(acc[i][j], SLOPE);
        }
        __syncthreads();
        {   // second layers: thread = (row, quarter of the hidden vector)
            constexpr int QW = C::HH / 4;
            const float* ar = sA + prow * C::LDA + q * QW;
            float sd = 0.f;
#pragma unroll
            for (int c = 0; c < QW; c += 4) {
                float4 a = ld4(ar + c), w = ld4(sW2 + q * QW + c);
                sd = fmaf(a.x, w.x, fmaf(a.y, w.y, fmaf(a.z, w.z, fmaf(a.w, w.w, sd))));
            }
            sd += __shfl_xor_sync(0xffffffffu, sd, 1);
            sd += __shfl_xor_sync(0xffffffffu, sd, 2);
            const long long row = row0 + prow;
            if (q == 0 && row < R) {
                float z = sd + b_d;
                p[row] = (1.f / (1.f + expf(-z))) * (1.f - 2.f * D_EPS) + D_EPS;
            }
            if constexpr (HASG) {
                float av[QW];
#pragma unroll
                for (int c = 0; c < QW; c += 4) {
                    float4 a = ld4(ar + C::HH + c);
                    av[c] = a.x; av[c + 1] = a.y; av[c + 2] = a.z; av[c + 3] = a.w;
                }
                for (int g = 0; g < G; ++g) {
                    const float* wr = sW2 + (1 + g) * C::LDW2 + q * QW;
                    float sg = 0.f;
#pragma unroll
                    for (int c = 0; c < QW; c += 4) {
                        float4 w = ld4(wr + c);
                        sg = fmaf(av[c], w.x, fmaf(av[c + 1], w.y, fmaf(av[c + 2], w.z, fmaf(av[c + 3], w.w, sg))));
                    }
                    sg += __shfl_xor_sync(0xffffffffu, sg, 1);
                    sg += __shfl_xor_sync(0xffffffffu, sg, 2);
                    if (q == 0 && row < R) branch[(size_t)row * G + g] = sg + __ldg(bg2 + g);
                }
            }
        }
    }
}

template <int TOD, bool HASG>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
disc_heads_bwd_kernel(const float* __restrict__ pe, int n, int k, const float* __restrict__ base,
                      const float* __restrict__ soc0, const float* __restrict__ W1p, const float* __restrict__ Wd2,
                      const float* __restrict__ Wg2, int G, const float* __restrict__ p, const float* __restrict__ dp,
                      const float* __restrict__ dbranch, float* __restrict__ d_pe, float* __restrict__ d_soc0,
                      float* __restrict__ d_base) {
    using C = Cfg<TOD, HASG>;
    extern __shared__ __align__(16) float smem[];
    float* sW1 = smem;                          // [NZ][LDK]
    float* sPe = sW1 + C::NZ * LDK;             // [ROWS][LDK]
    float* sA = sPe + ROWS * LDK;               // [ROWS][LDA]  z, then dz
    float* sW2 = sA + ROWS * C::LDA;            // [1 + G][LDW2]
    float* sDo = sW2 + (1 + GMAX) * C::LDW2;    // [ROWS][1 + GMAX]  d logit, d branch
    constexpr int LDO = 1 + GMAX;
    stage_weights<TOD, HASG>(sW1, sW2, W1p, Wd2, Wg2, G);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = (warp & 3) * 8 + (lane & 7);
    const int rl = (warp >> 2) * 32 + (lane >> 3);
    const int d_kq = threadIdx.x & 7, d_r0 = threadIdx.x >> 3;
    RowMap rm;
    rm.k = k; rm.n = n; rm.apt = ROWS / k;                 // k <= ROWS (checked by the caller)
    const long long n_tiles = ((long long)n + rm.apt - 1) / rm.apt;

    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        rm.agent0 = tile * rm.apt;
        __syncthreads();
        stage_tile_grouped<TOD, HASG>(sA, sPe, pe, base, soc0, rm);
        for (int idx = threadIdx.x; idx < ROWS * LDO; idx += MGGAN_THREADS) {
            int r = idx / LDO, c = idx - r * LDO;
            long long row; int ag_, smp_;
            float v = 0.f;
            if (rm.map(r, row, ag_, smp_)) {
                if (c == 0) {
                    if (dp != nullptr) {
                        float s = (__ldg(p + row) - D_EPS) / (1.f - 2.f * D_EPS);
                        v = __ldg(dp + row) * (1.f - 2.f * D_EPS) * s * (1.f - s);
                    }
                } else if (HASG && c - 1 < G && dbranch != nullptr) {
                    v = __ldg(dbranch + (size_t)row * G + c - 1);
                }
            }
            sDo[idx] = v;
        }
        __syncthreads();
        {
            float acc[8][C::TO];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < C::TO; ++j) acc[i][j] = sA[(rl + 4 * i) * C::LDA + u + 32 * j];
            tile_rowdot<8, C::TO, KP>(acc, sPe, LDK, rl, 4, sW1, LDK, u, 32);
            // dz = (upstream through the second layer) * LeakyReLU'(z)
            float t[8][C::TO];
            float wd[TOD];
#pragma unroll
            for (int j = 0; j < TOD; ++j) wd[j] = sW2[u + 32 * j];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float dl = sDo[(rl + 4 * i) * LDO];
#pragma unroll
                for (int j = 0; j < TOD; ++j) t[i][j] = dl * wd[j];
                if constexpr (HASG) {
#pragma unroll
                    for (int j = 0; j < TOD; ++j) t[i][TOD + j] = 0.f;
                }
            }
            if constexpr (HASG) {
                for (int g = 0; g < G; ++g) {
                    float wg[TOD];
#pragma unroll
                    for (int j = 0; j < TOD; ++j) wg[j] = sW2[(1 + g) * C::LDW2 + u + 32 * j];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float db = sDo[(rl + 4 * i) * LDO + 1 + g];
#pragma unroll
                        for (int j = 0; j < TOD; ++j) t[i][TOD + j] = fmaf(db, wg[j], t[i][TOD + j]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int r = rl + 4 * i;
                long long row; int ag, smp;
                const bool ok = rm.map(r, row, ag, smp);
#pragma unroll
                for (int j = 0; j < C::TO; ++j) {
                    const float dz = ok ? t[i][j] * (acc[i][j] > 0.f ? 1.f : SLOPE) : 0.f;
                    sA[r * C::LDA + u + 32 * j] = dz;
                    if (ok && smp == 0) d_soc0[(size_t)ag * C::NZ + u + 32 * j] = dz;
                }
            }
        }
        __syncthreads();
        {
            float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
            tile_dgrad<2, C::NZ>(acc, sA, C::LDA, d_r0, 32, sW1, LDK, d_kq * 4);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                long long row; int ag, smp;
                if (rm.map(d_r0 + 32 * h, row, ag, smp))
                    *reinterpret_cast<float4*>(d_pe + (size_t)row * KP + d_kq * 4) = make_float4(acc[h][0], acc[h][1], acc[h][2], acc[h][3]);
            }
        }
        if (d_base != nullptr) {     // d base[agent] = sum over the agent's k rows of the tile (plain stores: each agent is in one tile)
            for (int idx = threadIdx.x; idx < rm.apt * C::NZ; idx += MGGAN_THREADS) {
                const int a = idx / C::NZ, c = idx - a * C::NZ;
                const long long ag = rm.agent0 + a;
                if (ag < n) {
                    float sum = 0.f;
                    for (int q = 0; q < k; ++q) sum += sA[(a * k + q) * C::LDA + c];
                    d_base[(size_t)ag * C::NZ + c] = sum;
                }
            }
        }
    }
}

template <int TOD, bool HASG>
size_t fwd_smem(int G) {
    using C = Cfg<TOD, HASG>;
    return sizeof(float) * (C::NZ * LDK + ROWS * LDK + ROWS * C::LDA + (1 + G) * C::LDW2);
}
template <int TOD, bool HASG>
size_t bwd_smem() {
    using C = Cfg<TOD, HASG>;
    return sizeof(float) * (C::NZ * LDK + ROWS * LDK + ROWS * C::LDA + (1 + GMAX) * C::LDW2 + ROWS * (1 + GMAX));
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

int grid_for(long long R) {
    long long tiles = (R + ROWS - 1) / ROWS;
    long long cap = (long long)sm_count() * 2;
    return (int)(tiles < cap ? tiles : cap);
}

template <int TOD, bool HASG>
int launch_fwd(const float* pe, int n, int k, const float* base, const float* soc0, const float* W1p, const float* Wd2,
               const float* bd2, const float* Wg2, const float* bg2, int G, float* p, float* branch, cudaStream_t s) {
    size_t sm = fwd_smem<TOD, HASG>(G);
    cudaFuncSetAttribute(disc_heads_fwd_kernel<TOD, HASG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    disc_heads_fwd_kernel<TOD, HASG><<<grid_for((long long)n * k), MGGAN_THREADS, sm, s>>>(pe, n, k, base, soc0, W1p, Wd2, bd2,
                                                                                          Wg2, bg2, G, p, branch);
    return mggan_check_launch("disc_heads_fwd");
}
template <int TOD, bool HASG>
int launch_bwd(const float* pe, int n, int k, const float* base, const float* soc0, const float* W1p, const float* Wd2,
               const float* Wg2, int G, const float* p, const float* dp, const float* dbranch, float* d_pe, float* d_soc0,
               float* d_base, cudaStream_t s) {
    size_t sm = bwd_smem<TOD, HASG>();
    cudaFuncSetAttribute(disc_heads_bwd_kernel<TOD, HASG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    const long long tiles = ((long long)n + ROWS / k - 1) / (ROWS / k), cap = (long long)sm_count() * 2;
    disc_heads_bwd_kernel<TOD, HASG><<<(int)(tiles < cap ? tiles : cap), MGGAN_THREADS, sm, s>>>(pe, n, k, base, soc0, W1p, Wd2, Wg2, G, p,
                                                                                          dp, dbranch, d_pe, d_soc0, d_base);
    return mggan_check_launch("disc_heads_bwd");
}

}  // namespace

extern "C" int mggan_disc_heads_fwd(const float* pe, int n, int k, int HH, const float* base, const float* soc0,
                                    const float* W1p, const float* Wd2, const float* bd2, const float* Wg2,
                                    const float* bg2, int G, float* p, float* branch, cudaStream_t stream) {
    MGGAN_REQUIRE(n >= 0 && k >= 1 && (HH == 64 || HH == 96), "mggan_disc_heads_fwd: n=%d k=%d HH=%d (HH must be 64 or 96)", n, k, HH);
    MGGAN_REQUIRE(G >= 0 && G <= GMAX, "mggan_disc_heads_fwd: %d generators (max %d)", G, GMAX);
    MGGAN_REQUIRE((G == 0) == (branch == nullptr), "mggan_disc_heads_fwd: branch output and G must both be set or both empty");
    if (n == 0) return MGGAN_OK;
    if (HH == 96) {
        if (G > 0) return launch_fwd<3, true>(pe, n, k, base, soc0, W1p, Wd2, bd2, Wg2, bg2, G, p, branch, stream);
        return launch_fwd<3, false>(pe, n, k, base, soc0, W1p, Wd2, bd2, Wg2, bg2, G, p, branch, stream);
    }
    if (G > 0) return launch_fwd<2, true>(pe, n, k, base, soc0, W1p, Wd2, bd2, Wg2, bg2, G, p, branch, stream);
    return launch_fwd<2, false>(pe, n, k, base, soc0, W1p, Wd2, bd2, Wg2, bg2, G, p, branch, stream);
}

extern "C" int mggan_disc_heads_bwd(const float* pe, int n, int k, int HH, const float* base, const float* soc0,
                                    const float* W1p, const float* Wd2, const float* Wg2, int G, const float* p,
                                    const float* dp, const float* dbranch, float* d_pe, float* d_soc0, float* d_base,
                                    cudaStream_t stream) {
    MGGAN_REQUIRE(n >= 0 && k >= 1 && (HH == 64 || HH == 96), "mggan_disc_heads_bwd: n=%d k=%d HH=%d (HH must be 64 or 96)", n, k, HH);
    MGGAN_REQUIRE(G >= 0 && G <= GMAX, "mggan_disc_heads_bwd: %d generators (max %d)", G, GMAX);
    MGGAN_REQUIRE(d_pe != nullptr && d_soc0 != nullptr, "mggan_disc_heads_bwd: d_pe and d_soc0 are required");
    MGGAN_REQUIRE(k <= ROWS, "mggan_disc_heads_bwd: %d samples per agent (the backward tiles hold at most %d)", k, ROWS);
    if (n == 0) return MGGAN_OK;
    if (HH == 96) {
        if (G > 0) return launch_bwd<3, true>(pe, n, k, base, soc0, W1p, Wd2, Wg2, G, p, dp, dbranch, d_pe, d_soc0, d_base, stream);
        return launch_bwd<3, false>(pe, n, k, base, soc0, W1p, Wd2, Wg2, G, p, dp, dbranch, d_pe, d_soc0, d_base, stream);
    }
    if (G > 0) return launch_bwd<2, true>(pe, n, k, base, soc0, W1p, Wd2, Wg2, G, p, dp, dbranch, d_pe, d_soc0, d_base, stream);
    return launch_bwd<2, false>(pe, n, k, base, soc0, W1p, Wd2, Wg2, G, p, dp, dbranch, d_pe, d_soc0, d_base, stream);
}
