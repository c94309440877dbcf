// Fused trajectory-encoder LSTM (reference: mggan/model/modules/common_modules.py:24-66,
// TrajectoryEncoder = Linear(2,E) embedding + 1-layer nn.LSTM, returns h_T).
//
// The embedding is folded into the input projection on the host side (Wx = W_ih W_e, 4H x 2;
// b = W_ih b_e + b_ih + b_hh), so one step is  gates = Wx x_t + b + W_hh h_{t-1}  with the
// PyTorch gate order [i; f; g; o].  One persistent CTA owns a tile of rows (agents) for the
// whole sequence: W_hh stays in shared memory, h ping-pongs through shared memory, c lives in
// registers.  Forward and backward both run their per-step contractions on the tensor pipe
// (warp-level 3 x TF32 products, common.cuh).  Forward optionally saves (i, f, g, o, c, tanh c)
// per step for the backward, which walks the sequence in reverse and keeps the weight-gradient
// fragments in registers until the end (one atomicAdd per element per CTA).
#include "common.cuh"

namespace {

template <int H>
struct EncCfg {
    static constexpr int UG = H / 8;          // warps along hidden units
    static constexpr int RG = 8 / UG;         // warps along rows
    static constexpr int ROWS = RG * 32;      // rows per CTA tile
};

// Forward.  gates[row][4H] = (h_{t-1} | x_t | 1) . (W_hh | Wx | b)^T is one warp-level 3 x TF32 product per step
// (contraction H + 8: the input projection and the bias ride in three extra columns).  A warp owns 8 hidden units --
// the n-tiles of their four gates -- for two 16-row m-tiles, so the cell update is local to the thread that holds the
// C fragments: c stays in registers, h goes to the other half of the ping-pong tile as float2.  Both operand tiles
// have row stride 12 mod 32 (A[row g][k t] and B[gate g][k t] on banks 12 g + t).
template <int H>
struct EncFwd {
    static constexpr int ROWS = EncCfg<H>::ROWS;     // 32 (H = 64) or 64 (H = 32): two m-tiles per warp either way
    static constexpr int G4 = 4 * H;
    static constexpr int UG = H / 8;                 // unit groups
    static constexpr int LDW = H + 12;               // (W_hh | Wx0 Wx1 b 0 0 0 0 0) per gate row
    static constexpr int LDH = H + 12;               // (h | x0 x1 1 0 0 0 0 0) per row
    static constexpr size_t SMEM = sizeof(float) * (G4 * LDW + 2 * ROWS * LDH);
};

template <int H>
__global__ void __launch_bounds__(MGGAN_THREADS, 2)
lstm_enc_fwd_kernel(const float* __restrict__ x, int T, int N, const float* __restrict__ Wx,
                    const float* __restrict__ b, const float* __restrict__ Whh, float* __restrict__ hT,
                    float* __restrict__ acts) {
    using C = EncFwd<H>;
    constexpr int ROWS = C::ROWS, G4 = C::G4, UG = C::UG, LDW = C::LDW, LDH = C::LDH;
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                              // [4H][LDW]
    float* sH = sW + G4 * LDW;                     // [2][ROWS][LDH]
    for (int i = threadIdx.x; i < G4 * H; i += MGGAN_THREADS) sW[(i / H) * LDW + (i % H)] = __ldg(Whh + i);
    for (int o = threadIdx.x; o < G4; o += MGGAN_THREADS) {
        float* e = sW + o * LDW + H;
        e[0] = __ldg(Wx + 2 * o); e[1] = __ldg(Wx + 2 * o + 1); e[2] = __ldg(b + o);
        e[3] = 0.f; e[4] = 0.f; e[5] = 0.f; e[6] = 0.f; e[7] = 0.f;
    }
    for (int r = threadIdx.x; r < 2 * ROWS; r += MGGAN_THREADS) {
        float* e = sH + r * LDH + H;
        e[2] = 1.f; e[3] = 0.f; e[4] = 0.f; e[5] = 0.f; e[6] = 0.f; e[7] = 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int ug = warp % UG, m0 = (warp / UG) * 32;           // this warp's units 8 ug .. + 7, rows m0 .. m0 + 31
    const int u = ug * 8 + 2 * t4;                             // the C fragments hold units u, u + 1

    const int n_tiles = (N + ROWS - 1) / ROWS;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * ROWS;
        float c[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int e = 0; e < 4; ++e) c[i][e] = 0.f;
        __syncthreads();                           // weights staged / previous tile done with sH
        if (threadIdx.x < ROWS) {
            const int row = row0 + threadIdx.x;
            const float2 xv = row < N ? __ldg(reinterpret_cast<const float2*>(x) + row) : make_float2(0.f, 0.f);
            *reinterpret_cast<float2*>(sH + threadIdx.x * LDH + H) = xv;
        }
        __syncthreads();
        for (int t = 0; t < T; ++t) {
            const float* hcur = sH + (t & 1) * ROWS * LDH;
            float* hnext = sH + ((t + 1) & 1) * ROWS * LDH;
            float acc[2][4][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int e = 0; e < 4; ++e) acc[i][q][e] = 0.f;
            auto kstep = [&](int k0) {
                uint32_t ah[2][4], al[2][4];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float* ha = hcur + (m0 + 16 * i + g8) * LDH + k0 + t4;
                    tf32_split(ha[0], ah[i][0], al[i][0]);
                    tf32_split(ha[8 * LDH], ah[i][1], al[i][1]);
                    tf32_split(ha[4], ah[i][2], al[i][2]);
                    tf32_split(ha[8 * LDH + 4], ah[i][3], al[i][3]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float* wb = sW + (q * H + ug * 8 + g8) * LDW + k0 + t4;
                    uint32_t b0h, b0l, b1h, b1l;
                    tf32_split(wb[0], b0h, b0l);
                    tf32_split(wb[4], b1h, b1l);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        mma_tf32_16x8x8(acc[i][q], ah[i], b0h, b1h);
                        mma_tf32_16x8x8(acc[i][q], al[i], b0h, b1h);
                        mma_tf32_16x8x8(acc[i][q], ah[i], b0l, b1l);
                    }
                }
            };
            kstep(H);                              // the (x | 1) columns; h_{-1} = 0: the first step has nothing else
            if (t > 0) {
#pragma unroll 2
                for (int k0 = 0; k0 < H; k0 += 8) kstep(k0);
            }
            // cell update on the C fragments: element e = 2 half + unit (rows g8 + 8 half, units u + {0, 1})
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int r = m0 + 16 * i + g8 + 8 * half, row = row0 + r;
                    float ig[2], fg[2], gg[2], og[2], tc[2], hh[2];
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        ig[e] = sigmoidf_(acc[i][0][2 * half + e]);
                        fg[e] = sigmoidf_(acc[i][1][2 * half + e]);
                        gg[e] = tanhf_(acc[i][2][2 * half + e]);
                        og[e] = sigmoidf_(acc[i][3][2 * half + e]);
                        c[i][2 * half + e] = fmaf(fg[e], c[i][2 * half + e], ig[e] * gg[e]);
                        tc[e] = tanhf_(c[i][2 * half + e]);
                        hh[e] = og[e] * tc[e];
                    }
                    *reinterpret_cast<float2*>(hnext + r * LDH + u) = make_float2(hh[0], hh[1]);
                    if (row < N) {
                        if (acts != nullptr) {
                            float* a = acts + ((size_t)t * N + row) * (6 * H) + u;
                            *reinterpret_cast<float2*>(a) = make_float2(ig[0], ig[1]);
                            *reinterpret_cast<float2*>(a + H) = make_float2(fg[0], fg[1]);
                            *reinterpret_cast<float2*>(a + 2 * H) = make_float2(gg[0], gg[1]);
                            *reinterpret_cast<float2*>(a + 3 * H) = make_float2(og[0], og[1]);
                            *reinterpret_cast<float2*>(a + 4 * H) = make_float2(c[i][2 * half], c[i][2 * half + 1]);
                            *reinterpret_cast<float2*>(a + 5 * H) = make_float2(tc[0], tc[1]);
                        }
                        if (t == T - 1) *reinterpret_cast<float2*>(hT + (size_t)row * H + u) = make_float2(hh[0], hh[1]);
                    }
                }
            if (t + 1 < T && threadIdx.x < ROWS) {           // x_{t+1} next to h_t
                const int row = row0 + threadIdx.x;
                const float2 xv = row < N ? __ldg(reinterpret_cast<const float2*>(x) + (size_t)(t + 1) * N + row) : make_float2(0.f, 0.f);
                *reinterpret_cast<float2*>(hnext + threadIdx.x * LDH + H) = xv;
            }
            __syncthreads();
        }
    }
}

// Backward.  Per step the two contractions -- dh_{t-1} = dgates . W_hh (rows x 4H x H) and
// dW_hh += dgates^T . h_{t-1} (4H x rows x H) -- run as warp-level 3 x TF32 products (common.cuh) over a 64-row tile:
// the FP32 register tiles they replace re-read every operand once per 4 x 4 outputs and kept the FMA pipe 31 % busy at
// 177 registers (one CTA of 8 warps per SM).  The h_{t-1} tile carries three more columns (x0, x1, 1), so the same
// weight-gradient product also yields dWx and db (one more n-tile instead of a 64-row scalar loop per gate).
// Shared-memory strides are chosen per fragment pattern: the dgates tile is read as A[row g][gate t] by the input
// gradient (stride = 4 mod 32: banks 4 g + t) and as A[gate g][row 2 t] by the weight gradient, whose contraction index
// is permuted (fragment k = t <-> row 2 t, k = t + 4 <-> row 2 t + 1: banks 8 t + g); W_hh rows are 8 mod 32 apart
// (B[gate t][unit g]: 8 t + g) and the h tile's 12 mod 32 (B[row 2 t][unit g]: 24 t + g).
template <int H>
struct EncBwd {
    static constexpr int ROWS = 64;
    static constexpr int G4 = 4 * H;
    static constexpr int LDW = H + 8;
    static constexpr int LDG = G4 + 4;
    static constexpr int LDH = H + 12;               // columns H .. H + 7: x0, x1, 1, 0, 0, 0, 0, 0
    static constexpr int LDD = H + 8;                // dh tile: float2 stores of the C fragments on 8 g + 2 t
    static constexpr int NTW = H / 8 + 1;            // n-tiles of the weight gradient (the last one: dWx | db)
    static constexpr int MTW = G4 / 16 / 8;          // its m-tiles (16 gates each) per warp: 1 or 2
    static constexpr int NTD = H / 16;               // n-tiles of the input gradient per warp: 2 or 4
    static constexpr int UH = H / 32;                // 32-unit groups
    static constexpr int ITEMS = ROWS * H / MGGAN_THREADS;     // (row, unit) pairs per thread in the cell backward
    static constexpr size_t SMEM = sizeof(float) * (G4 * LDW + ROWS * LDG + ROWS * LDH + ROWS * LDD);
};

template <int H>
__global__ void __launch_bounds__(MGGAN_THREADS, H == 32 ? 2 : 1)
lstm_enc_bwd_kernel(const float* __restrict__ x, int T, int N, const float* __restrict__ Whh,
                    const float* __restrict__ acts, const float* __restrict__ dhT, float* __restrict__ dWx,
                    float* __restrict__ db, float* __restrict__ dWhh) {
    using C = EncBwd<H>;
    constexpr int ROWS = C::ROWS, G4 = C::G4, LDW = C::LDW, LDG = C::LDG, LDH = C::LDH, LDD = C::LDD;
    constexpr int NTW = C::NTW, MTW = C::MTW, NTD = C::NTD, UH = C::UH, ITEMS = C::ITEMS;
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                      // [4H][LDW]
    float* sG = sW + G4 * LDW;             // [ROWS][LDG]   gate pre-activation gradients
    float* sHp = sG + ROWS * LDG;          // [ROWS][LDH]   h_{t-1} | x_t | 1
    float* sDh = sHp + ROWS * LDH;         // [ROWS][LDD]   dL/dh_t
    for (int i = threadIdx.x; i < G4 * H; i += MGGAN_THREADS) sW[(i / H) * LDW + (i % H)] = __ldg(Whh + i);
    if (threadIdx.x < ROWS) {
        float* e = sHp + threadIdx.x * LDH + H;
        e[0] = 0.f; e[1] = 0.f; e[2] = 1.f; e[3] = 0.f; e[4] = 0.f; e[5] = 0.f; e[6] = 0.f; e[7] = 0.f;
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g8 = lane >> 2, t4 = lane & 3;
    const int md = warp & 3, nd0 = (warp >> 2) * NTD;          // input gradient: m-tile, first n-tile of this warp
    float wacc[MTW][NTW][4];
#pragma unroll
    for (int i = 0; i < MTW; ++i)
#pragma unroll
        for (int j = 0; j < NTW; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) wacc[i][j][e] = 0.f;

    const int n_tiles = (N + ROWS - 1) / ROWS;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * ROWS;
        float dc[ITEMS];
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) dc[j] = 0.f;
        __syncthreads();                   // weights staged / the previous tile's products are done with the tiles
#pragma unroll
        for (int j = 0; j < ITEMS; ++j) {
            const int u = lane + 32 * (j % UH), r = warp + 8 * (j / UH), row = row0 + r;
            sDh[r * LDD + u] = row < N ? __ldg(dhT + (size_t)row * H + u) : 0.f;
        }
        __syncthreads();
        for (int t = T - 1; t >= 0; --t) {
            // ---- phase 1: cell backward, thread = (unit lane + 32 ., rows warp + 8 .): every load is a full line.
            // Loads are unconditional (row clamped, result masked) and issued for 4 pairs at a time.
#pragma unroll
            for (int b4 = 0; b4 < ITEMS; b4 += 4) {
                float v[4][5], pv[4][3];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = b4 + jj;
                    const int u = lane + 32 * (j % UH), row = min(row0 + warp + 8 * (j / UH), N - 1);
                    const float* a = acts + ((size_t)t * N + row) * (6 * H) + u;
                    v[jj][0] = __ldg(a); v[jj][1] = __ldg(a + H); v[jj][2] = __ldg(a + 2 * H); v[jj][3] = __ldg(a + 3 * H);
                    v[jj][4] = __ldg(a + 5 * H);
                    if (t > 0) {
                        const float* ap = a - (size_t)N * (6 * H);
                        pv[jj][0] = __ldg(ap + 4 * H); pv[jj][1] = __ldg(ap + 3 * H); pv[jj][2] = __ldg(ap + 5 * H);
                    }
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = b4 + jj;
                    const int u = lane + 32 * (j % UH), r = warp + 8 * (j / UH);
                    const bool valid = row0 + r < N;
                    const float ig = v[jj][0], fg = v[jj][1], gg = v[jj][2], og = v[jj][3], tc = v[jj][4];
                    const float cp = t > 0 ? pv[jj][0] : 0.f;
                    const float hp = t > 0 ? pv[jj][1] * pv[jj][2] : 0.f;
                    const float dh = sDh[r * LDD + u];
                    const float dcc = fmaf(dh * og, 1.f - tc * tc, dc[j]);
                    const float dao = dh * tc * og * (1.f - og);
                    const float dai = dcc * gg * ig * (1.f - ig);
                    const float dag = dcc * ig * (1.f - gg * gg);
                    const float daf = dcc * cp * fg * (1.f - fg);
                    dc[j] = valid ? dcc * fg : 0.f;
                    float* gr = sG + r * LDG + u;
                    gr[0] = valid ? dai : 0.f; gr[H] = valid ? daf : 0.f; gr[2 * H] = valid ? dag : 0.f; gr[3 * H] = valid ? dao : 0.f;
                    sHp[r * LDH + u] = valid ? hp : 0.f;
                }
            }
            if (threadIdx.x < ROWS) {
                const int row = row0 + threadIdx.x;
                const float2 xv = row < N ? __ldg(reinterpret_cast<const float2*>(x) + (size_t)t * N + row) : make_float2(0.f, 0.f);
                *reinterpret_cast<float2*>(sHp + threadIdx.x * LDH + H) = xv;
            }
            __syncthreads();
            // ---- phase 2a: dh_{t-1}[row][n] = sum_gate dgates[row][gate] W_hh[gate][n]
            if (t > 0) {
                float dacc[NTD][4];
#pragma unroll
                for (int j = 0; j < NTD; ++j)
#pragma unroll
                    for (int e = 0; e < 4; ++e) dacc[j][e] = 0.f;
                const float* ga = sG + (md * 16 + g8) * LDG + t4;
#pragma unroll 2
                for (int k0 = 0; k0 < G4; k0 += 8) {
                    uint32_t ah[4], al[4];
                    tf32_split(ga[k0], ah[0], al[0]);
                    tf32_split(ga[k0 + 8 * LDG], ah[1], al[1]);
                    tf32_split(ga[k0 + 4], ah[2], al[2]);
                    tf32_split(ga[k0 + 8 * LDG + 4], ah[3], al[3]);
                    const float* wb = sW + (k0 + t4) * LDW + nd0 * 8 + g8;
#pragma unroll
                    for (int j = 0; j < NTD; ++j) mma_3xtf32(dacc[j], ah, al, wb[8 * j], wb[8 * j + 4 * LDW]);
                }
                // dh of step t was consumed in phase 1 (before the barrier above): safe to overwrite
#pragma unroll
                for (int j = 0; j < NTD; ++j) {
                    float* d = sDh + (md * 16 + g8) * LDD + (nd0 + j) * 8 + 2 * t4;
                    *reinterpret_cast<float2*>(d) = make_float2(dacc[j][0], dacc[j][1]);
                    *reinterpret_cast<float2*>(d + 8 * LDD) = make_float2(dacc[j][2], dacc[j][3]);
                }
            }
            // ---- phase 2b: dW[gate][n] += sum_row dgates[row][gate] (h_{t-1} | x | 1)[row][n]
#pragma unroll 1
            for (int r0 = 2 * t4; r0 < ROWS; r0 += 8) {
                uint32_t ah[MTW][4], al[MTW][4];
#pragma unroll
                for (int i = 0; i < MTW; ++i) {
                    const float* ga = sG + r0 * LDG + (warp * MTW + i) * 16 + g8;
                    tf32_split(ga[0], ah[i][0], al[i][0]);
                    tf32_split(ga[8], ah[i][1], al[i][1]);
                    tf32_split(ga[LDG], ah[i][2], al[i][2]);
                    tf32_split(ga[LDG + 8], ah[i][3], al[i][3]);
                }
                const float* hb = sHp + r0 * LDH + g8;
#pragma unroll
                for (int j = 0; j < NTW; ++j) {
                    uint32_t b0h, b0l, b1h, b1l;
                    tf32_split(hb[8 * j], b0h, b0l);
                    tf32_split(hb[8 * j + LDH], b1h, b1l);
#pragma unroll
                    for (int i = 0; i < MTW; ++i) {
                        mma_tf32_16x8x8(wacc[i][j], ah[i], b0h, b1h);
                        mma_tf32_16x8x8(wacc[i][j], al[i], b0h, b1h);
                        mma_tf32_16x8x8(wacc[i][j], ah[i], b0l, b1l);
                    }
                }
            }
            __syncthreads();
        }
    }
    // C fragments -> global: c0 (gate m0 + g, col n0 + 2t), c1 (., + 1), c2 / c3: gate + 8
#pragma unroll
    for (int i = 0; i < MTW; ++i) {
        const int m = (warp * MTW + i) * 16 + g8;
#pragma unroll
        for (int j = 0; j < NTW - 1; ++j) {
            float* d = dWhh + (size_t)m * H + 8 * j + 2 * t4;
            atomicAdd(d, wacc[i][j][0]);
            atomicAdd(d + 1, wacc[i][j][1]);
            atomicAdd(d + 8 * H, wacc[i][j][2]);
            atomicAdd(d + 8 * H + 1, wacc[i][j][3]);
        }
        if (t4 == 0) {                                   // columns 0, 1 of the last n-tile: dWx
            atomicAdd(dWx + 2 * m, wacc[i][NTW - 1][0]);
            atomicAdd(dWx + 2 * m + 1, wacc[i][NTW - 1][1]);
            atomicAdd(dWx + 2 * (m + 8), wacc[i][NTW - 1][2]);
            atomicAdd(dWx + 2 * (m + 8) + 1, wacc[i][NTW - 1][3]);
        } else if (t4 == 1) {                            // column 2: db
            atomicAdd(db + m, wacc[i][NTW - 1][0]);
            atomicAdd(db + m + 8, wacc[i][NTW - 1][2]);
        }
    }
}

template <int H>
size_t enc_fwd_smem() { return EncFwd<H>::SMEM; }
template <int H>
size_t enc_bwd_smem() { return EncBwd<H>::SMEM; }

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
    size_t sm = enc_bwd_smem<H>();
    cudaFuncSetAttribute(lstm_enc_bwd_kernel<H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int tiles = (N + EncBwd<H>::ROWS - 1) / EncBwd<H>::ROWS;
    int per_sm = H == 32 ? 2 : 1;
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
