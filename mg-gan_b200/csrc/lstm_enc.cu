// Fused trajectory-encoder LSTM (reference: mggan/model/modules/common_modules.py:24-66,
// TrajectoryEncoder = Linear(2,E) embedding + 1-layer nn.LSTM, returns h_T).
//
// The embedding is folded into the input projection on the host side (Wx = W_ih W_e, 4H x 2;
// b = W_ih b_e + b_ih + b_hh), so one step is  gates = Wx x_t + b + W_hh h_{t-1}  with the
// PyTorch gate order [i; f; g; o].  One persistent CTA owns a tile of rows (agents) for the
// whole sequence: W_hh stays in shared memory, h ping-pongs through shared memory, c lives in
// registers.  Forward optionally saves (i, f, g, o, c, tanh c) per step for the backward,
// which walks the sequence in reverse and keeps the weight-gradient tiles in registers until
// the end (one atomicAdd per element per CTA).
#include "common.cuh"

namespace {

template <int H>
struct EncCfg {
    static constexpr int UG = H / 8;          // warps along hidden units
    static constexpr int RG = 8 / UG;         // warps along rows
    static constexpr int ROWS = RG * 32;      // rows per CTA tile
    static constexpr int LDH = H + 4;
    static constexpr int LDG = 4 * H + 4;
};

template <int H>
__global__ void __launch_bounds__(MGGAN_THREADS)
lstm_enc_fwd_kernel(const float* __restrict__ x, int T, int N, const float* __restrict__ Wx,
                    const float* __restrict__ b, const float* __restrict__ Whh, float* __restrict__ hT,
                    float* __restrict__ acts) {
    using C = EncCfg<H>;
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                              // [4H][LDH]
    float* sH = sW + 4 * H * C::LDH;               // [2][ROWS][LDH]
    stage_matrix(sW, C::LDH, Whh, 4 * H, H);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = (warp % C::UG) * 8 + (lane & 7);
    const int rl = (warp / C::UG) * 32 + (lane >> 3);
    float wx0[4], wx1[4], bb[4];
#pragma unroll
    for (int g = 0; g < 4; ++g) {
        wx0[g] = __ldg(Wx + (g * H + u) * 2);
        wx1[g] = __ldg(Wx + (g * H + u) * 2 + 1);
        bb[g] = __ldg(b + g * H + u);
    }
    const int n_tiles = (N + C::ROWS - 1) / C::ROWS;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * C::ROWS;
        float c[8], h[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i] = 0.f, h[i] = 0.f;
        __syncthreads();                           // previous tile done with sH / weights staged
        for (int t = 0; t < T; ++t) {
            const float* hcur = sH + (t & 1) * C::ROWS * C::LDH;
            float* hnext = sH + ((t + 1) & 1) * C::ROWS * C::LDH;
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int row = row0 + rl + 4 * i;
                float2 xv = row < N ? __ldg(reinterpret_cast<const float2*>(x) + (size_t)t * N + row) : make_float2(0.f, 0.f);
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[i][g] = fmaf(wx0[g], xv.x, fmaf(wx1[g], xv.y, bb[g]));
            }
            if (t > 0) tile_rowdot<8, 4, H>(acc, hcur, C::LDH, rl, 4, sW, C::LDH, u, H);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float ig = sigmoidf_(acc[i][0]), fg = sigmoidf_(acc[i][1]);
                float gg = tanhf_(acc[i][2]), og = sigmoidf_(acc[i][3]);
                c[i] = fmaf(fg, c[i], ig * gg);
                float tc = tanhf_(c[i]);
                h[i] = og * tc;
                hnext[(rl + 4 * i) * C::LDH + u] = h[i];
                int row = row0 + rl + 4 * i;
                if (acts != nullptr && row < N) {
                    float* a = acts + ((size_t)t * N + row) * (6 * H) + u;
                    a[0] = ig; a[H] = fg; a[2 * H] = gg; a[3 * H] = og; a[4 * H] = c[i]; a[5 * H] = tc;
                }
            }
            __syncthreads();
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = row0 + rl + 4 * i;
            if (row < N) hT[(size_t)row * H + u] = h[i];
        }
    }
}

template <int H>
__global__ void __launch_bounds__(MGGAN_THREADS)
lstm_enc_bwd_kernel(const float* __restrict__ x, int T, int N, const float* __restrict__ Whh,
                    const float* __restrict__ acts, const float* __restrict__ dhT, float* __restrict__ dWx,
                    float* __restrict__ db, float* __restrict__ dWhh) {
    using C = EncCfg<H>;
    constexpr int ROWS = C::ROWS, LDH = C::LDH, LDG = C::LDG;
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                      // [4H][LDH]
    float* sG = sW + 4 * H * LDH;          // [ROWS][LDG]   gate pre-activation gradients
    float* sHp = sG + ROWS * LDG;          // [ROWS][LDH]   h_{t-1}
    float* sDh = sHp + ROWS * LDH;         // [ROWS][LDH]   dL/dh_t
    float* sX = sDh + ROWS * LDH;          // [ROWS][2]
    stage_matrix(sW, LDH, Whh, 4 * H, H);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = (warp % C::UG) * 8 + (lane & 7);
    const int rl = (warp / C::UG) * 32 + (lane >> 3);
    // dgrad mapping: thread -> (k-quad, 2 rows)
    constexpr int KQ = H / 4;
    const int d_kq = threadIdx.x % KQ, d_r0 = threadIdx.x / KQ;
    constexpr int D_RS = MGGAN_THREADS / KQ;   // ROWS == 2 * D_RS
    static_assert(ROWS == 2 * D_RS, "dgrad mapping");
    // wgrad mapping: thread -> NB 4x4 blocks of dWhh
    constexpr int NB = (4 * H * H / 16) / MGGAN_THREADS;
    float wacc[NB][4][4];
#pragma unroll
    for (int j = 0; j < NB; ++j)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int bq = 0; bq < 4; ++bq) wacc[j][a][bq] = 0.f;
    float ax0 = 0.f, ax1 = 0.f, ab = 0.f;          // dWx[o][0..1], db[o] for o = threadIdx.x < 4H

    const int n_tiles = (N + ROWS - 1) / ROWS;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * ROWS;
        float dc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dc[i] = 0.f;
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = rl + 4 * i, row = row0 + r;
            sDh[r * LDH + u] = row < N ? __ldg(dhT + (size_t)row * H + u) : 0.f;
        }
        __syncthreads();
        for (int t = T - 1; t >= 0; --t) {
            // ---- phase 1: cell backward.  Activation loads are unconditional (row clamped, result masked) and issued
            // for 4 rows at a time so that their latencies overlap instead of adding up.
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float v[4][5], pv[4][3];
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int row = min(row0 + rl + 4 * (half * 4 + ii), N - 1);
                    const float* a = acts + ((size_t)t * N + row) * (6 * H) + u;
                    v[ii][0] = __ldg(a); v[ii][1] = __ldg(a + H); v[ii][2] = __ldg(a + 2 * H); v[ii][3] = __ldg(a + 3 * H);
                    v[ii][4] = __ldg(a + 5 * H);
                    if (t > 0) {
                        const float* ap = a - (size_t)N * (6 * H);
                        pv[ii][0] = __ldg(ap + 4 * H); pv[ii][1] = __ldg(ap + 3 * H); pv[ii][2] = __ldg(ap + 5 * H);
                    }
                }
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int i = half * 4 + ii;
                    const int r = rl + 4 * i;
                    const bool valid = row0 + r < N;
                    const float ig = v[ii][0], fg = v[ii][1], gg = v[ii][2], og = v[ii][3], tc = v[ii][4];
                    const float cp = t > 0 ? pv[ii][0] : 0.f;
                    const float hp = t > 0 ? pv[ii][1] * pv[ii][2] : 0.f;
                    const float dh = sDh[r * LDH + u];
                    const float dcc = fmaf(dh * og, 1.f - tc * tc, dc[i]);
                    const float dao = dh * tc * og * (1.f - og);
                    const float dai = dcc * gg * ig * (1.f - ig);
                    const float dag = dcc * ig * (1.f - gg * gg);
                    const float daf = dcc * cp * fg * (1.f - fg);
                    dc[i] = valid ? dcc * fg : 0.f;
                    sG[r * LDG + u] = valid ? dai : 0.f; sG[r * LDG + H + u] = valid ? daf : 0.f;
                    sG[r * LDG + 2 * H + u] = valid ? dag : 0.f; sG[r * LDG + 3 * H + u] = valid ? dao : 0.f;
                    sHp[r * LDH + u] = valid ? hp : 0.f;
                }
            }
            for (int i = threadIdx.x; i < ROWS; i += MGGAN_THREADS) {
                int row = row0 + i;
                float2 xv = row < N ? __ldg(reinterpret_cast<const float2*>(x) + (size_t)t * N + row) : make_float2(0.f, 0.f);
                sX[2 * i] = xv.x; sX[2 * i + 1] = xv.y;
            }
            __syncthreads();
            // ---- phase 2: dh_{t-1} = W_hh^T dgates ; weight-gradient tiles
            if (t > 0) {
                float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                tile_dgrad<2, 4 * H>(acc, sG, LDG, d_r0, D_RS, sW, LDH, d_kq * 4);
#pragma unroll
                for (int j = 0; j < NB; ++j) {
                    int oq = (H == 32) ? (warp & 3) * 8 + (lane & 7) : warp * 8 + (lane & 7);
                    int kq = (H == 32) ? (warp >> 2) * 4 + (lane >> 3) : (lane >> 3) + 4 * j;
                    tile_wgrad<ROWS>(wacc[j], sG, LDG, oq * 4, sHp, LDH, kq * 4);
                }
                // sDh of step t was consumed in phase 1 (before the barrier above): safe to overwrite
#pragma unroll
                for (int i = 0; i < 2; ++i)
                    st4(sDh + (d_r0 + i * D_RS) * LDH + d_kq * 4, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
            }
            if (threadIdx.x < 4 * H) {
                int o = threadIdx.x;
#pragma unroll 4
                for (int r = 0; r < ROWS; ++r) {
                    float g = sG[r * LDG + o];
                    ax0 = fmaf(g, sX[2 * r], ax0);
                    ax1 = fmaf(g, sX[2 * r + 1], ax1);
                    ab += g;
                }
            }
            __syncthreads();
        }
    }
#pragma unroll
    for (int j = 0; j < NB; ++j) {
        int oq = (H == 32) ? (warp & 3) * 8 + (lane & 7) : warp * 8 + (lane & 7);
        int kq = (H == 32) ? (warp >> 2) * 4 + (lane >> 3) : (lane >> 3) + 4 * j;
        atomic_block44(dWhh, H, oq * 4, kq * 4, wacc[j]);
    }
    if (threadIdx.x < 4 * H) {
        atomicAdd(dWx + threadIdx.x * 2, ax0);
        atomicAdd(dWx + threadIdx.x * 2 + 1, ax1);
        atomicAdd(db + threadIdx.x, ab);
    }
}

template <int H>
size_t enc_fwd_smem() { using C = EncCfg<H>; return sizeof(float) * (4 * H * C::LDH + 2 * C::ROWS * C::LDH); }
template <int H>
size_t enc_bwd_smem() {
    using C = EncCfg<H>;
    return sizeof(float) * (4 * H * C::LDH + C::ROWS * C::LDG + 2 * C::ROWS * C::LDH + 2 * C::ROWS);
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

template <int H>
int launch_fwd(const float* x, int T, int N, const float* Wx, const float* b, const float* Whh, float* hT, float* acts,
               cudaStream_t s) {
    using C = EncCfg<H>;
    size_t sm = enc_fwd_smem<H>();
    cudaFuncSetAttribute(lstm_enc_fwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int tiles = (N + C::ROWS - 1) / C::ROWS;
    int grid = tiles < sm_count() * 2 ? tiles : sm_count() * 2;
    lstm_enc_fwd_kernel<H><<<grid, MGGAN_THREADS, sm, s>>>(x, T, N, Wx, b, Whh, hT, acts);
    return mggan_check_launch("lstm_enc_fwd");
}
template <int H>
int launch_bwd(const float* x, int T, int N, const float* Whh, const float* acts, const float* dhT, float* dWx, float* db,
               float* dWhh, cudaStream_t s) {
    using C = EncCfg<H>;
    size_t sm = enc_bwd_smem<H>();
    cudaFuncSetAttribute(lstm_enc_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int tiles = (N + C::ROWS - 1) / C::ROWS;
    int per_sm = sm <= 110 * 1024 ? 2 : 1;
    int grid = tiles < sm_count() * per_sm ? tiles : sm_count() * per_sm;
    lstm_enc_bwd_kernel<H><<<grid, MGGAN_THREADS, sm, s>>>(x, T, N, Whh, acts, dhT, dWx, db, dWhh);
    return mggan_check_launch("lstm_enc_bwd");
}

}  // namespace

extern "C" int mggan_lstm_seq_fwd(const float* x, int T, int N, int H, const float* Wx, const float* b,
                                  const float* Whh, float* hT, float* acts, cudaStream_t stream) {
    MGGAN_REQUIRE(H == 32 || H == 64, "mggan_lstm_seq_fwd: hidden size %d not built (32 or 64)", H);
    MGGAN_REQUIRE(T >= 1 && N >= 0, "mggan_lstm_seq_fwd: bad T=%d N=%d", T, N);
    if (N == 0) return MGGAN_OK;
    return H == 32 ? launch_fwd<32>(x, T, N, Wx, b, Whh, hT, acts, stream)
                   : launch_fwd<64>(x, T, N, Wx, b, Whh, hT, acts, stream);
}

extern "C" int mggan_lstm_seq_bwd(const float* x, int T, int N, int H, const float* Whh, const float* acts,
                                  const float* dhT, float* dWx, float* db, float* dWhh, cudaStream_t stream) {
    MGGAN_REQUIRE(H == 32 || H == 64, "mggan_lstm_seq_bwd: hidden size %d not built (32 or 64)", H);
    MGGAN_REQUIRE(T >= 1 && N >= 0 && acts != nullptr, "mggan_lstm_seq_bwd: bad arguments");
    if (N == 0) return MGGAN_OK;
    return H == 32 ? launch_bwd<32>(x, T, N, Whh, acts, dhT, dWx, db, dWhh, stream)
                   : launch_bwd<64>(x, T, N, Whh, acts, dhT, dWx, db, dWhh, stream);
}
