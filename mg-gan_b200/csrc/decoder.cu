// Fused multi-generator trajectory decoder (reference: RelativeDecoder.forward
// mggan/model/modules/common_modules.py:97-131, called once per generator and per sample
// multiplicity by MultiGenerator.forward_all mggan/model/modules/standard.py:227-265, followed
// by the (sample, generator) gather of standard.py:190-214).
//
// The reference decodes G * M * N sequences with 12 * G separate LSTM-step launches and then
// gathers k * N of them.  Here the *selected* sequences are the unit of work: rows are
// sequences sorted by generator (select.cu builds the order), a CTA owns 64 rows of one
// generator for all pred_len autoregressive steps, and results are written straight to their
// (t, sample, agent) slot.  Per step and row:
//     gates = Wx dxdy_{t-1} + b + W_hh h_{t-1}      (Wx = W_ih W_s folded on the host, 128 x 2)
//     (i, f, g, o) -> c_t, h_t
//     u = LReLU_0.01(W_1h h_t + [W_1s social + b_1])   (social half hoisted out of the time loop)
//     dxdy_t = W_2 u + b_2 ;  xy_t = xy_{t-1} + dxdy_t
// h_0 = A[agent] + Wz z   where A = enc_h_to_dec_h applied to the per-agent encoding (hoisted;
// computed once per agent by the caller) and Wz is the noise block of that layer.
#include "common.cuh"

namespace {

constexpr int H = 32;        // decoder hidden size (config.decoder_h_dim default)
constexpr int M1 = 16;       // hidden2pos mid width = H / 2
constexpr int ROWS = 64;
// Row strides = 8 mod 32: the (4 rows x 8 units) scalar accesses of a warp (rows lane>>3, units lane&7) then fall on
// 32 distinct banks (with H + 4 the rows were 4 banks apart and every such access was a 2-way conflict).
constexpr int LDH = H + 8;   // 40
constexpr int LDG = 4 * H + 8;
constexpr int LDU = M1 + 4;  // 20
constexpr int ZMAX = 16;

struct DecWeights {
    const float* Wz;    // (H, Z)            shared by all generators
    const float* Wx;    // (G, 4H, 2)
    const float* b;     // (G, 4H)
    const float* Whh;   // (G, 4H, H)
    const float* W1h;   // (G, M1, H)
    const float* W1s;   // (G, M1, H)
    const float* b1;    // (G, M1)
    const float* W2;    // (G, 2, M1)
    const float* b2;    // (G, 2)
};

struct DecGrads {
    float* dWz; float* dWx; float* db; float* dWhh; float* dW1h; float* dW1s; float* db1; float* dW2; float* db2;
    float* dA;        // (n_agents, H)
    float* dsocial;   // (n_agents, H)
};

struct DecSeq {
    const int* tile_gen;    // (n_tiles) generator of each 64-row tile, -1 = unused tile
    const int* seq_agent;   // (n_tiles*64) agent index of the row, -1 = padding
    const int* seq_noise;   // row into noise (.., Z)
    const int* seq_out;     // output column
};

__global__ void __launch_bounds__(MGGAN_THREADS, 2)
decoder_fwd_kernel(DecSeq sq, int n_tiles, const float* __restrict__ A, const float* __restrict__ social,
                   const float* __restrict__ last_xy, const float* __restrict__ last_dxdy,
                   const float* __restrict__ noise, int Z, DecWeights w, int T, int n_cols,
                   float* __restrict__ out_abs, float* __restrict__ out_rel, float* __restrict__ acts,
                   float* __restrict__ u1save, float* __restrict__ h0save) {
    extern __shared__ __align__(16) float smem[];
    float* sW = smem;                       // [4H][LDH]
    float* sW1 = sW + 4 * H * LDH;          // [M1][LDH]   W1h
    float* sW1s = sW1 + M1 * LDH;           // [M1][LDH]   W1s
    float* sH = sW1s + M1 * LDH;            // [2][ROWS][LDH]
    float* sBs = sH + 2 * ROWS * LDH;       // [ROWS][LDU]
    float* sD = sBs + ROWS * LDU;           // [ROWS][2]
    float* sWz = sD + ROWS * 2;             // [H][ZMAX+1]
    __shared__ int sAgent[ROWS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = (warp & 3) * 8 + (lane & 7);
    const int rl = (warp >> 2) * 32 + (lane >> 3);
    const int prow = threadIdx.x >> 2, mq = threadIdx.x & 3;     // hidden2pos mapping
    const size_t Rpad = (size_t)n_tiles * ROWS;

    for (int i = threadIdx.x; i < H * Z; i += MGGAN_THREADS) sWz[(i / Z) * (ZMAX + 1) + (i % Z)] = __ldg(w.Wz + i);

    const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per, t_end = min(n_tiles, t_begin + per);
    int cur_g = -1;
    float wx0[4], wx1[4], bb[4], w2a[4], w2b[4], b2a = 0.f, b2b = 0.f;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int g = sq.tile_gen[tile];
        if (g < 0) continue;
        __syncthreads();                       // previous tile finished with all shared buffers
        if (g != cur_g) {
            cur_g = g;
            stage_matrix(sW, LDH, w.Whh + (size_t)g * 4 * H * H, 4 * H, H);
            stage_matrix(sW1, LDH, w.W1h + (size_t)g * M1 * H, M1, H);
            stage_matrix(sW1s, LDH, w.W1s + (size_t)g * M1 * H, M1, H);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                wx0[q] = __ldg(w.Wx + ((size_t)g * 4 * H + q * H + u) * 2);
                wx1[q] = __ldg(w.Wx + ((size_t)g * 4 * H + q * H + u) * 2 + 1);
                bb[q] = __ldg(w.b + (size_t)g * 4 * H + q * H + u);
                w2a[q] = __ldg(w.W2 + (size_t)g * 2 * M1 + mq + 4 * q);
                w2b[q] = __ldg(w.W2 + (size_t)g * 2 * M1 + M1 + mq + 4 * q);
            }
            b2a = __ldg(w.b2 + g * 2);
            b2b = __ldg(w.b2 + g * 2 + 1);
        }
        const int row0 = tile * ROWS;
        if (threadIdx.x < ROWS) sAgent[threadIdx.x] = sq.seq_agent[row0 + threadIdx.x];
        __syncthreads();
        // ---- per-row constants: Bs = b1 + W1s social ; xy, dxdy of the last observation
        const int pag = sAgent[prow];
        float xy0 = 0.f, xy1 = 0.f;
        {
            float bs[1][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) bs[0][j] = __ldg(w.b1 + (size_t)g * M1 + mq + 4 * j);
            if (pag >= 0) {
                const float* sp = social + (size_t)pag * H;
#pragma unroll
                for (int k = 0; k < H; k += 4) {
                    float4 s4 = __ldg(reinterpret_cast<const float4*>(sp + k));
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float4 w4 = ld4(sW1s + (mq + 4 * j) * LDH + k);
                        bs[0][j] = fmaf(s4.x, w4.x, fmaf(s4.y, w4.y, fmaf(s4.z, w4.z, fmaf(s4.w, w4.w, bs[0][j]))));
                    }
                }
                xy0 = __ldg(last_xy + (size_t)pag * 2);
                xy1 = __ldg(last_xy + (size_t)pag * 2 + 1);
                if (mq == 0) {
                    sD[prow * 2] = __ldg(last_dxdy + (size_t)pag * 2);
                    sD[prow * 2 + 1] = __ldg(last_dxdy + (size_t)pag * 2 + 1);
                }
            } else if (mq == 0) {
                sD[prow * 2] = 0.f; sD[prow * 2 + 1] = 0.f;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) sBs[prow * LDU + mq + 4 * j] = bs[0][j];
        }
        // ---- h0 = A[agent] + Wz z ; c0 = 0
        float c[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = rl + 4 * i, ag = sAgent[r];
            float h0 = 0.f;
            if (ag >= 0) {
                h0 = __ldg(A + (size_t)ag * H + u);
                const float* zp = noise + (size_t)sq.seq_noise[row0 + r] * Z;
                for (int z = 0; z < Z; ++z) h0 = fmaf(sWz[u * (ZMAX + 1) + z], __ldg(zp + z), h0);
                if (h0save != nullptr) h0save[dec_h0_off(row0 + r, u)] = h0;
            }
            sH[r * LDH + u] = h0;
            c[i] = 0.f;
        }
        const int pcol = pag >= 0 ? sq.seq_out[row0 + prow] : -1;
        __syncthreads();
        for (int t = 0; t < T; ++t) {
            const float* hcur = sH + (t & 1) * ROWS * LDH;
            float* hnext = sH + ((t + 1) & 1) * ROWS * LDH;
            float acc[8][4];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float d0 = sD[(rl + 4 * i) * 2], d1 = sD[(rl + 4 * i) * 2 + 1];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][q] = fmaf(wx0[q], d0, fmaf(wx1[q], d1, bb[q]));
            }
            tile_rowdot<8, 4, H>(acc, hcur, LDH, rl, 4, sW, LDH, u, H);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int r = rl + 4 * i;
                float ig = sigmoidf_(acc[i][0]), fg = sigmoidf_(acc[i][1]);
                float gg = tanhf_(acc[i][2]), og = sigmoidf_(acc[i][3]);
                c[i] = fmaf(fg, c[i], ig * gg);
                float tc = tanhf_(c[i]);
                hnext[r * LDH + u] = og * tc;
                if (acts != nullptr && sAgent[r] >= 0) {
                    // layout: dec_acts_off (common.cuh); a warp's float2 stores cover 64-byte segments
                    *reinterpret_cast<float2*>(acts + dec_acts_off(Rpad >> 7, t, row0 + r, u)) = make_float2(og * tc, c[i]);
                }
            }
            __syncthreads();
            // ---- hidden2pos on h_t
            {
                float up[1][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) up[0][j] = sBs[prow * LDU + mq + 4 * j];
                tile_rowdot<1, 4, H>(up, hnext, LDH, prow, 0, sW1, LDH, mq, 4);
                float d0 = 0.f, d1 = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float a = lrelu_(up[0][j], 0.01f);
                    d0 = fmaf(w2a[j], a, d0);
                    d1 = fmaf(w2b[j], a, d1);
                }
                d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
                d0 += __shfl_xor_sync(0xffffffffu, d0, 2); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
                d0 += b2a; d1 += b2b;
                xy0 += d0; xy1 += d1;
                if (pcol >= 0) {
                    if (u1save != nullptr) {
                        float* us = u1save + dec_u1_off(Rpad >> 7, t, row0 + prow, mq);
#pragma unroll
                        for (int j = 0; j < 4; ++j) us[j * 512] = up[0][j];       // m = mq + 4 j
                    }
                    if (mq == 0) {
                        size_t o = ((size_t)t * n_cols + pcol) * 2;
                        *reinterpret_cast<float2*>(out_rel + o) = make_float2(d0, d1);
                        *reinterpret_cast<float2*>(out_abs + o) = make_float2(xy0, xy1);
                    }
                }
                if (mq == 0) { sD[prow * 2] = d0; sD[prow * 2 + 1] = d1; }
            }
            __syncthreads();
        }
    }
}

// Backward.  Inputs: the forward's saved recurrent state (h_t, c_t per step), d_abs / d_rel (either may be null).
// Per step and 64-row tile:
//   phase 0  hidden2pos backward (thread = row x 4 mid units); the tile's h_{t-1}, c_{t-1} (16 contiguous 1 KB chunks,
//            float4 loads issued before the phase's arithmetic) -> shared memory, with (x0, x1, 1) in the pad columns
//   phase G  the gate pre-activations are RECOMPUTED on the tensor pipe:  Z = [h_{t-1} | x | 1] [W_hh | Wx | b]^T
//            (64 x 128, K = 40) -- round 1 re-read the six saved gate / cell values per unit and step (4.1 GB per
//            generator step at 21 % of the HBM peak, long-scoreboard stalls 28 % of the kernel) where this reads two
//   phase 1  LSTM cell backward (thread = 8 rows x 1 unit): gates from Z, c_t = f c_{t-1} + i g, d(gates) in place of Z
//   phase 2  tile products:
//   dh_{t-1} = dG W_hh (64 x 32, K = 128)          dW_hh += dG^T h_{t-1} (128 x 32, K = 64 rows)
//       -- these two are 80 % of the step's MACs and made the kernel shared-memory-wavefront bound (lsu 86 %) as FP32
//          register tiles; they run as warp-level 3 x TF32 mma.sync products (common.cuh), 5-6x fewer wavefronts per MAC
//   (dWx | db) += dG^T (x | 1)   via the pad columns 32..34 of the h_{t-1} tile, rows split over the k-quad lanes
//   dW1h += dU^T h_t (16 x 32)   rows split over the 8 warps
//   d(dxdy_{t-1}) = dG Wx        thread = (row, quarter of the 128 gates), Wx kept transposed (rows 32, 33 of the W_hh^T tile)
// All weight-gradient tiles stay in registers across steps and tiles of one generator and are flushed with one
// atomicAdd per element per CTA.
constexpr int KG = H + 8;     // contraction length of the gate recompute: 32 units, x0, x1, 1, 5 zero columns

__global__ void __launch_bounds__(MGGAN_THREADS, 2)
decoder_bwd_kernel(DecSeq sq, int n_tiles, const float* __restrict__ social, const float* __restrict__ last_dxdy,
                   const float* __restrict__ noise, int Z, DecWeights w, int T, int n_cols,
                   const float* __restrict__ out_rel, const float* __restrict__ acts,
                   const float* __restrict__ u1save, const float* __restrict__ h0save,
                   const float* __restrict__ d_abs, const float* __restrict__ d_rel, DecGrads gr) {
    extern __shared__ __align__(16) float smem[];
    float* sWT = smem;                      // [KG][LDG]   rows 0..31 W_hh transposed (unit-major): B operand of dh = dG W_hh and of the
                                            //             gate recompute; rows 32, 33 Wx transposed, row 34 b, rows 35..39 zero
    float* sW1 = sWT + KG * LDG;            // [M1][LDH]   W1s (epilogue)
    float* sG = sW1 + M1 * LDH;             // [ROWS][LDG] gate pre-activation gradients
    float* sHp = sG + ROWS * LDG;           // [ROWS][LDH] h_{t-1} | x0 x1 1 0
    float* sHt = sHp + ROWS * LDH;          // [ROWS][LDH] h_t   (social tile after the loop)
    float* sDh = sHt + ROWS * LDH;          // [ROWS][LDH] dL/dh from the later step
    float* sDu = sDh + ROWS * LDH;          // [ROWS][LDU] d(hidden2pos.0 pre-activation)
    float* sCp = sDu + ROWS * LDU;          // [ROWS][LDH] c_{t-1}
    float* sWxT = sWT + H * LDG;            //             Wx transposed = rows 32, 33 of sWT (row stride LDG)
    float* sZ = sCp + ROWS * LDH;           // [ROWS][ZMAX] noise rows of the tile
    float* sW1sAcc = sZ + ROWS * ZMAX;      // [M1][H]     dW1s of the current generator (per-tile shared-memory adds)
    float* sW1h = sW1sAcc + M1 * H;         // [M1][LDH]   W1h
    float* sW2 = sW1h + M1 * LDH;           // [2][M1]     hidden2pos.2
    __shared__ int sAgent[ROWS];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int u = (warp & 3) * 8 + (lane & 7);
    const int rl = (warp >> 2) * 32 + (lane >> 3);
    const int prow = threadIdx.x >> 2, mq = threadIdx.x & 3;
    const int e_mq = lane & 3, e_kq = lane >> 2;                    // dW1h / dW1s 4x4 block, rows split over warps
    const int z_u = threadIdx.x & 31, z_g = threadIdx.x >> 5;       // dWz: unit, noise-column group (z_g, z_g + 8)
    const size_t Rpad = (size_t)n_tiles * ROWS;

    float wacc[5][4], w1acc[4];         // wacc[0..3]: dW_hh n-tiles; wacc[4]: the pad columns (x0, x1, 1) = (dWx | db); w1acc: dW1h (warps 0-3)
    float aw2a[4], aw2b[4], ab2a, ab2b, ab1[4], awz0, awz1;
    auto zero_acc = [&]() {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) wacc[a][b] = 0.f;
            wacc[4][a] = 0.f; w1acc[a] = 0.f;
            aw2a[a] = aw2b[a] = ab1[a] = 0.f;
        }
        ab2a = ab2b = awz0 = awz1 = 0.f;
        for (int i = threadIdx.x; i < M1 * H; i += MGGAN_THREADS) sW1sAcc[i] = 0.f;     // read again only after a barrier
    };
    auto flush = [&](int g) {
        {   // wacc[j] = C fragment of n-tile j: c0, c1 -> gate m0 + 2g, units 8t + j, 8t + 4 + j; c2, c3 -> gate m0 + 2g + 1
            float* dst = gr.dWhh + (size_t)g * 4 * H * H + (size_t)(warp * 16 + 2 * (lane >> 2)) * H + 8 * (lane & 3);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                atomicAdd(dst + j, wacc[j][0]); atomicAdd(dst + 4 + j, wacc[j][1]);
                atomicAdd(dst + H + j, wacc[j][2]); atomicAdd(dst + H + 4 + j, wacc[j][3]);
            }
        }
        {   // wacc[4]: c0, c1 -> gate m0 + 2g, pad columns 2t, 2t + 1 (0: x0, 1: x1, 2: the constant 1); c2, c3 -> gate m0 + 2g + 1
            const size_t o = (size_t)g * 4 * H + warp * 16 + 2 * (lane >> 2);
            if ((lane & 3) == 0) {
                atomicAdd(gr.dWx + o * 2, wacc[4][0]); atomicAdd(gr.dWx + o * 2 + 1, wacc[4][1]);
                atomicAdd(gr.dWx + o * 2 + 2, wacc[4][2]); atomicAdd(gr.dWx + o * 2 + 3, wacc[4][3]);
            } else if ((lane & 3) == 1) {
                atomicAdd(gr.db + o, wacc[4][0]); atomicAdd(gr.db + o + 1, wacc[4][2]);
            }
        }
        if (warp < 4) {   // w1acc: c0, c1 -> mid unit g, hidden units 8 warp + 2t, + 1; c2, c3 -> mid unit g + 8
            float* dst = gr.dW1h + (size_t)g * M1 * H + (size_t)(lane >> 2) * H + 8 * warp + 2 * (lane & 3);
            atomicAdd(dst, w1acc[0]); atomicAdd(dst + 1, w1acc[1]);
            atomicAdd(dst + 8 * H, w1acc[2]); atomicAdd(dst + 8 * H + 1, w1acc[3]);
        }
        for (int i = threadIdx.x; i < M1 * H; i += MGGAN_THREADS) atomicAdd(gr.dW1s + (size_t)g * M1 * H + i, sW1sAcc[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            atomicAdd(gr.dW2 + (size_t)g * 2 * M1 + mq + 4 * j, aw2a[j]);
            atomicAdd(gr.dW2 + (size_t)g * 2 * M1 + M1 + mq + 4 * j, aw2b[j]);
            atomicAdd(gr.db1 + (size_t)g * M1 + mq + 4 * j, ab1[j]);
        }
        if (mq == 0) {
            atomicAdd(gr.db2 + g * 2, ab2a);
            atomicAdd(gr.db2 + g * 2 + 1, ab2b);
        }
        if (z_g < Z) atomicAdd(gr.dWz + z_u * Z + z_g, awz0);
        if (z_g + 8 < Z) atomicAdd(gr.dWz + z_u * Z + z_g + 8, awz1);
    };
    zero_acc();
    for (int i = threadIdx.x; i < (KG - H - 3) * LDG; i += MGGAN_THREADS) sWT[(H + 3) * LDG + i] = 0.f;     // zero rows of the K extension
    for (int i = threadIdx.x; i < ROWS * (LDH - H); i += MGGAN_THREADS) sHp[(i / (LDH - H)) * LDH + H + i % (LDH - H)] = 0.f;

    const int per = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int t_begin = blockIdx.x * per, t_end = min(n_tiles, t_begin + per);
    int cur_g = -1;
    for (int tile = t_begin; tile < t_end; ++tile) {
        const int g = sq.tile_gen[tile];
        if (g < 0) continue;
        __syncthreads();
        if (g != cur_g) {
            if (cur_g >= 0) { flush(cur_g); __syncthreads(); zero_acc(); }
            cur_g = g;
            for (int i = threadIdx.x; i < 4 * H * H; i += MGGAN_THREADS)
                sWT[(i & (H - 1)) * LDG + (i >> 5)] = __ldg(w.Whh + (size_t)g * 4 * H * H + i);
            for (int i = threadIdx.x; i < 4 * H * 2; i += MGGAN_THREADS)
                sWxT[(i & 1) * LDG + (i >> 1)] = __ldg(w.Wx + (size_t)g * 4 * H * 2 + i);
            for (int i = threadIdx.x; i < 4 * H; i += MGGAN_THREADS) sWT[(H + 2) * LDG + i] = __ldg(w.b + (size_t)g * 4 * H + i);
            stage_matrix(sW1h, LDH, w.W1h + (size_t)g * M1 * H, M1, H);
            if (threadIdx.x < 2 * M1) sW2[threadIdx.x] = __ldg(w.W2 + (size_t)g * 2 * M1 + threadIdx.x);
        }
        const int row0 = tile * ROWS;
        if (threadIdx.x < ROWS) sAgent[threadIdx.x] = sq.seq_agent[row0 + threadIdx.x];
        float dhreg[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};     // dL/dh of the warp's block from the later step
        __syncthreads();
        const int pag = sAgent[prow];
        const int pcol = pag >= 0 ? sq.seq_out[row0 + prow] : -1;
        {   // noise rows of the tile (for dWz in the epilogue)
            const int r = threadIdx.x >> 2;
            const int ag = sAgent[r];
            const float* zp = noise + (size_t)(ag >= 0 ? sq.seq_noise[row0 + r] : 0) * Z;
            for (int z = threadIdx.x & 3; z < ZMAX; z += 4) sZ[r * ZMAX + z] = (ag >= 0 && z < Z) ? __ldg(zp + z) : 0.f;
        }
        // output gradients of the row: the values of step t - 1 are requested during step t (they are 8-byte gathers by output
        // column whose latency sat in front of every step: 6 % of the kernel's stall samples)
        float2 g_abs = make_float2(0.f, 0.f), g_rel = make_float2(0.f, 0.f);
        if (pcol >= 0) {
            const size_t o = ((size_t)(T - 1) * n_cols + pcol) * 2;
            if (d_abs != nullptr) g_abs = __ldg(reinterpret_cast<const float2*>(d_abs + o));
            if (d_rel != nullptr) g_rel = __ldg(reinterpret_cast<const float2*>(d_rel + o));
        }
        float dxy0 = 0.f, dxy1 = 0.f;       // sum_{tau >= t} d_abs[tau]
        float dn0 = 0.f, dn1 = 0.f;         // gradient reaching dxdy_t through step t+1's input
        float dbs[4] = {0.f, 0.f, 0.f, 0.f};
        float dc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) dc[i] = 0.f;

        for (int t = T - 1; t >= 0; --t) {
            // ---- phase 0: hidden2pos backward (thread = row x 4 mid units) on values requested one step ago.  They are consumed
            // BEFORE this step's loads are issued: a register whose load shares a scoreboard with younger loads waits for those too
            // (the gradient gather showed 8 % long-scoreboard stalls although it had been in flight for a whole step).
            float dr0 = dn0, dr1 = dn1;
            if (pcol >= 0) {
                dxy0 += g_abs.x; dxy1 += g_abs.y;
                dr0 += dxy0 + g_rel.x; dr1 += dxy1 + g_rel.y;
            }
            // this step's loads: the tile's recurrent state of the previous step (4 float4 per thread = (h, c) of two units;
            // t = 0: h_0 from its own buffer, c_0 = 0), the row's hidden2pos pre-activations and step input, and the next
            // step's output gradients
            float4 hc[4];
            float2 xv = make_float2(0.f, 0.f);
            float u1p[4] = {0.f, 0.f, 0.f, 0.f};
            {
                const size_t ns = Rpad >> 7;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int idx = threadIdx.x + q * MGGAN_THREADS, r = idx & (ROWS - 1), j = idx >> 6;
                    if (t > 0) hc[q] = __ldg(reinterpret_cast<const float4*>(acts + dec_acts_off(ns, t - 1, row0 + r, 2 * j)));
                    else if (q < 2) hc[q] = __ldg(reinterpret_cast<const float4*>(h0save + dec_h0_off(row0 + r, 4 * j)));
                }
                if (pcol >= 0) {
                    const float* us = u1save + dec_u1_off(ns, t, row0 + prow, mq);
#pragma unroll
                    for (int j = 0; j < 4; ++j) u1p[j] = us[j * 512];                 // m = mq + 4 j
                    if (mq == 0)
                        xv = t > 0 ? __ldg(reinterpret_cast<const float2*>(out_rel + ((size_t)(t - 1) * n_cols + pcol) * 2))
                                   : __ldg(reinterpret_cast<const float2*>(last_dxdy + (size_t)pag * 2));
                    if (t > 0) {
                        const size_t o = ((size_t)(t - 1) * n_cols + pcol) * 2;
                        if (d_abs != nullptr) g_abs = __ldg(reinterpret_cast<const float2*>(d_abs + o));
                        if (d_rel != nullptr) g_rel = __ldg(reinterpret_cast<const float2*>(d_rel + o));
                    }
                }
            }
            float du[4] = {0.f, 0.f, 0.f, 0.f};
            if (pcol >= 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float w2a = sW2[mq + 4 * j], w2b = sW2[M1 + mq + 4 * j];
                    const float up = u1p[j];
                    const float a = lrelu_(up, 0.01f);
                    aw2a[j] = fmaf(dr0, a, aw2a[j]);
                    aw2b[j] = fmaf(dr1, a, aw2b[j]);
                    du[j] = (w2a * dr0 + w2b * dr1) * (up > 0.f ? 1.f : 0.01f);
                    dbs[j] += du[j];
                }
                if (mq == 0) { ab2a += dr0; ab2b += dr1; }
            }
            {
#pragma unroll
                for (int j = 0; j < 4; ++j) sDu[prow * LDU + mq + 4 * j] = du[j];
                // step input dxdy_{t-1} and the constant 1 go to the pad columns of the h_{t-1} tile
                if (mq == 0) st4(sHp + prow * LDH + H, make_float4(xv.x, xv.y, pcol >= 0 ? 1.f : 0.f, 0.f));
                // h_{t-1}, c_{t-1} -> shared memory (padding rows: zeros)
                if (t > 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int idx = threadIdx.x + q * MGGAN_THREADS, r = idx & (ROWS - 1), j = idx >> 6;
                        const bool ok = sAgent[r] >= 0;
                        *reinterpret_cast<float2*>(sHp + r * LDH + 2 * j) = ok ? make_float2(hc[q].x, hc[q].z) : make_float2(0.f, 0.f);
                        *reinterpret_cast<float2*>(sCp + r * LDH + 2 * j) = ok ? make_float2(hc[q].y, hc[q].w) : make_float2(0.f, 0.f);
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int idx = threadIdx.x + q * MGGAN_THREADS, r = idx & (ROWS - 1), j = idx >> 6;
                        if (q < 2) st4(sHp + r * LDH + 4 * j, sAgent[r] >= 0 ? hc[q] : make_float4(0.f, 0.f, 0.f, 0.f));
                        *reinterpret_cast<float2*>(sCp + r * LDH + 2 * j) = make_float2(0.f, 0.f);
                    }
                }
            }
            __syncthreads();
            // ---- phase G + 1: the warp owns rows m0 .. m0 + 15 x units n0 .. n0 + 15 (the same block whose dh it produced in
            // phase 2 of the later step, kept in registers).  Gate pre-activations Z = [h_{t-1} | x | 1] [W_hh | Wx | b]^T as
            // warp-level 3 x TF32 products: K = 40, B fragment (k = t, n = g) = sWT[(k0 + t) LDG + col + g] (row stride = 8 mod
            // 32: 32 distinct banks); the four n-tiles of a pass are the gates (i, f, g, o) of 8 units, so every lane ends up with
            // the complete gates of its (2 rows x 2 units) and the cell backward runs on the accumulators: no round trip of Z
            // through shared memory and no barrier between the product and the cell.
            {
                const int g8 = lane >> 2, t4 = lane & 3;
                const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 16;
                float hacc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};      // dU_t W1h: the hidden2pos path into h_t
#pragma unroll
                for (int k0 = 0; k0 < M1; k0 += 8) {
                    const float* pa = sDu + (m0 + g8) * LDU + k0 + t4;
                    uint32_t ah[4], al[4], bh[4], bl[4];
                    tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8 * LDU], ah[1], al[1]);
                    tf32_split(pa[4], ah[2], al[2]); tf32_split(pa[8 * LDU + 4], ah[3], al[3]);
                    const float* pb = sW1h + (k0 + t4) * LDH + n0 + g8;
                    tf32_split(pb[0], bh[0], bl[0]); tf32_split(pb[4 * LDH], bh[1], bl[1]);
                    tf32_split(pb[8], bh[2], bl[2]); tf32_split(pb[4 * LDH + 8], bh[3], bl[3]);
                    mma_tf32_16x8x8(hacc[0], ah, bh[0], bh[1]); mma_tf32_16x8x8(hacc[1], ah, bh[2], bh[3]);
                    mma_tf32_16x8x8(hacc[0], al, bh[0], bh[1]); mma_tf32_16x8x8(hacc[1], al, bh[2], bh[3]);
                    mma_tf32_16x8x8(hacc[0], ah, bl[0], bl[1]); mma_tf32_16x8x8(hacc[1], ah, bl[2], bl[3]);
                }
#pragma unroll
                for (int jp = 0; jp < 2; ++jp) {
                    const int u0 = n0 + 8 * jp;                 // units u0 .. u0 + 7 of this pass
                    float acc[4][4];                            // [gate][row g: units 2t, 2t + 1 | row g + 8: units 2t, 2t + 1]
#pragma unroll
                    for (int q = 0; q < 4; ++q) { acc[q][0] = 0.f; acc[q][1] = 0.f; acc[q][2] = 0.f; acc[q][3] = 0.f; }
#pragma unroll
                    for (int k0 = 0; k0 < KG; k0 += 8) {
                        const float* pa = sHp + (m0 + g8) * LDH + k0 + t4;
                        uint32_t ah[4], al[4], bh[4][2], bl[4][2];
                        tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8 * LDH], ah[1], al[1]);
                        tf32_split(pa[4], ah[2], al[2]); tf32_split(pa[8 * LDH + 4], ah[3], al[3]);
                        const float* pb = sWT + (k0 + t4) * LDG + u0 + g8;
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            tf32_split(pb[H * q], bh[q][0], bl[q][0]);
                            tf32_split(pb[4 * LDG + H * q], bh[q][1], bl[q][1]);
                        }
#pragma unroll
                        for (int q = 0; q < 4; ++q) mma_tf32_16x8x8(acc[q], ah, bh[q][0], bh[q][1]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) mma_tf32_16x8x8(acc[q], al, bh[q][0], bh[q][1]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) mma_tf32_16x8x8(acc[q], ah, bl[q][0], bl[q][1]);
                    }
                    // LSTM cell backward on the accumulators: c_t = f c_{t-1} + i g; gate gradients -> sG, h_t -> sHt
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        const int r = m0 + g8 + 8 * hf;
                        const bool valid = sAgent[r] >= 0;
                        const float2 cpv = *reinterpret_cast<const float2*>(sCp + r * LDH + u0 + 2 * t4);
                        float dgi[2], dgf[2], dgg[2], dgo[2], htv[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int ci = 2 * hf + e;
                            const float ig = sigmoidf_(acc[0][ci]), fg = sigmoidf_(acc[1][ci]);
                            const float gg = tanhf_(acc[2][ci]), og = sigmoidf_(acc[3][ci]);
                            const float cp = e ? cpv.y : cpv.x;
                            const float tc = tanhf_(fmaf(fg, cp, ig * gg));
                            const float dh = dhreg[jp][ci] + hacc[jp][ci];
                            const float dcc = fmaf(dh * og, 1.f - tc * tc, dc[4 * jp + ci]);
                            dc[4 * jp + ci] = valid ? dcc * fg : 0.f;
                            dgo[e] = valid ? dh * tc * og * (1.f - og) : 0.f;
                            dgi[e] = valid ? dcc * gg * ig * (1.f - ig) : 0.f;
                            dgg[e] = valid ? dcc * ig * (1.f - gg * gg) : 0.f;
                            dgf[e] = valid ? dcc * cp * fg * (1.f - fg) : 0.f;
                            htv[e] = valid ? og * tc : 0.f;
                        }
                        float* og_ = sG + r * LDG + u0 + 2 * t4;
                        *reinterpret_cast<float2*>(og_) = make_float2(dgi[0], dgi[1]);
                        *reinterpret_cast<float2*>(og_ + H) = make_float2(dgf[0], dgf[1]);
                        *reinterpret_cast<float2*>(og_ + 2 * H) = make_float2(dgg[0], dgg[1]);
                        *reinterpret_cast<float2*>(og_ + 3 * H) = make_float2(dgo[0], dgo[1]);
                        *reinterpret_cast<float2*>(sHt + r * LDH + u0 + 2 * t4) = make_float2(htv[0], htv[1]);
                    }
                }
            }
            __syncthreads();
            // ---- phase 2: tile products
            {
                // dh_{t-1} = dG W_hh: warp = 16 rows x 16 units, K = 128 gates.  The K index is permuted (logical k = t,
                // t+4 <-> gates k0 + 2t, k0 + 2t + 1) so that every fragment is one conflict-free LDS.64.
                const int g8 = lane >> 2, t4 = lane & 3;
                {
                    const int m0 = (warp & 3) * 16, n0 = (warp >> 2) * 16;
                    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
                    const float* pa0 = sG + (m0 + g8) * LDG + 2 * t4;
                    const float* pa1 = pa0 + 8 * LDG;
                    const float* pb0 = sWT + (n0 + g8) * LDG + 2 * t4;
                    const float* pb1 = pb0 + 8 * LDG;
#pragma unroll 4
                    for (int k0 = 0; k0 < 4 * H; k0 += 8) {
                        const float2 x0 = *reinterpret_cast<const float2*>(pa0 + k0);
                        const float2 x1 = *reinterpret_cast<const float2*>(pa1 + k0);
                        const float2 y0 = *reinterpret_cast<const float2*>(pb0 + k0);
                        const float2 y1 = *reinterpret_cast<const float2*>(pb1 + k0);
                        uint32_t ah[4], al[4], bh[4], bl[4];
                        tf32_split(x0.x, ah[0], al[0]); tf32_split(x1.x, ah[1], al[1]);
                        tf32_split(x0.y, ah[2], al[2]); tf32_split(x1.y, ah[3], al[3]);
                        tf32_split(y0.x, bh[0], bl[0]); tf32_split(y0.y, bh[1], bl[1]);
                        tf32_split(y1.x, bh[2], bl[2]); tf32_split(y1.y, bh[3], bl[3]);
                        // product-major over the two accumulators (back-to-back MMAs on one accumulator wait for each other)
                        mma_tf32_16x8x8(acc[0], ah, bh[0], bh[1]); mma_tf32_16x8x8(acc[1], ah, bh[2], bh[3]);
                        mma_tf32_16x8x8(acc[0], al, bh[0], bh[1]); mma_tf32_16x8x8(acc[1], al, bh[2], bh[3]);
                        mma_tf32_16x8x8(acc[0], ah, bl[0], bl[1]); mma_tf32_16x8x8(acc[1], ah, bl[2], bl[3]);
                    }
                    // dh_{t-1} of the warp's (rows, units) block stays in registers for the next step's cell backward
#pragma unroll
                    for (int j = 0; j < 2; ++j) { dhreg[j][0] = acc[j][0]; dhreg[j][1] = acc[j][1]; dhreg[j][2] = acc[j][2]; dhreg[j][3] = acc[j][3]; }
                }
                // dW_hh += dG^T h_{t-1}: warp = 16 gates x 32 units, K = 64 rows.  Rows of the A fragment are permuted
                // (logical g, g+8 <-> gates m0 + 2g, m0 + 2g + 1: one LDS.64) and so are the columns of B (n-tile j, logical
                // column c <-> unit 4c + j: one LDS.128 serves the four n-tiles); wacc[j] is the C fragment of n-tile j.
                {
                    const int m0 = warp * 16;
                    const float* pa = sG + t4 * LDG + m0 + 2 * g8;
                    const float* pb = sHp + t4 * LDH + 4 * g8;
#pragma unroll 2
                    for (int k0 = 0; k0 < ROWS; k0 += 8) {
                        const float2 x0 = *reinterpret_cast<const float2*>(pa + k0 * LDG);
                        const float2 x1 = *reinterpret_cast<const float2*>(pa + (k0 + 4) * LDG);
                        const float4 y0 = ld4(pb + k0 * LDH);
                        const float4 y1 = ld4(pb + (k0 + 4) * LDH);
                        uint32_t ah[4], al[4];
                        tf32_split(x0.x, ah[0], al[0]); tf32_split(x0.y, ah[1], al[1]);
                        tf32_split(x1.x, ah[2], al[2]); tf32_split(x1.y, ah[3], al[3]);
                        // five n-tiles (the fifth: (dWx | db) += dG^T (x0 x1 1) over the pad columns of the h_{t-1} tile), MMAs
                        // product-major over the five accumulators
                        const float b0[5] = {y0.x, y0.y, y0.z, y0.w, pb[k0 * LDH + H - 3 * g8]};
                        const float b1[5] = {y1.x, y1.y, y1.z, y1.w, pb[(k0 + 4) * LDH + H - 3 * g8]};
                        uint32_t bh[5][2], bl[5][2];
#pragma unroll
                        for (int j = 0; j < 5; ++j) { tf32_split(b0[j], bh[j][0], bl[j][0]); tf32_split(b1[j], bh[j][1], bl[j][1]); }
#pragma unroll
                        for (int j = 0; j < 5; ++j) mma_tf32_16x8x8(wacc[j], ah, bh[j][0], bh[j][1]);
#pragma unroll
                        for (int j = 0; j < 5; ++j) mma_tf32_16x8x8(wacc[j], al, bh[j][0], bh[j][1]);
#pragma unroll
                        for (int j = 0; j < 5; ++j) mma_tf32_16x8x8(wacc[j], ah, bl[j][0], bl[j][1]);
                    }
                }
            }
            // dW1h: rows [8 warp, 8 warp + 8) of dU^T h_t
            if (warp < 4) {   // dW1h += dU^T h_t (16 x 32, K = 64 rows): warp = hidden units 8 warp .. + 7; A (m, k = row) = sDu[row LDU + m]
                const int g8 = lane >> 2, t4 = lane & 3;
#pragma unroll 2
                for (int k0 = 0; k0 < ROWS; k0 += 8) {
                    const float* pa = sDu + (k0 + t4) * LDU + g8;
                    const float* pb = sHt + (k0 + t4) * LDH + 8 * warp + g8;
                    uint32_t ah[4], al[4], bh[2], bl[2];
                    tf32_split(pa[0], ah[0], al[0]); tf32_split(pa[8], ah[1], al[1]);
                    tf32_split(pa[4 * LDU], ah[2], al[2]); tf32_split(pa[4 * LDU + 8], ah[3], al[3]);
                    tf32_split(pb[0], bh[0], bl[0]); tf32_split(pb[4 * LDH], bh[1], bl[1]);
                    mma_tf32_16x8x8(w1acc, ah, bh[0], bh[1]);
                    mma_tf32_16x8x8(w1acc, al, bh[0], bh[1]);
                    mma_tf32_16x8x8(w1acc, ah, bl[0], bl[1]);
                }
            }
            {   // gradient wrt this step's input dxdy_{t-1}: dG Wx (row = prow; the 4 lanes of a row take interleaved gate
                // quads, so each load instruction of the quad is 64 contiguous bytes -- a contiguous quarter per lane put
                // all four on one bank: 4-way conflicts on 24 LDS.128 per step, 16 % of the kernel's shared wavefronts)
                float s0 = 0.f, s1 = 0.f;
#pragma unroll
                for (int oo = 0; oo < H; oo += 4) {
                    const int o = oo * 4 + mq * 4;
                    const float4 gq = ld4(sG + prow * LDG + o);
                    const float4 wa = ld4(sWxT + o), wb = ld4(sWxT + LDG + o);
                    s0 = fmaf(gq.x, wa.x, fmaf(gq.y, wa.y, fmaf(gq.z, wa.z, fmaf(gq.w, wa.w, s0))));
                    s1 = fmaf(gq.x, wb.x, fmaf(gq.y, wb.y, fmaf(gq.z, wb.z, fmaf(gq.w, wb.w, s1))));
                }
                s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 1);
                s0 += __shfl_xor_sync(0xffffffffu, s0, 2); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
                dn0 = s0; dn1 = s1;
            }
            __syncthreads();
        }
        // ---- epilogue: h0 and the hoisted social / b1 terms (dh_0 -> shared memory for the row-wise passes below)
        {
            const int g8 = lane >> 2, t4 = lane & 3, m0 = (warp & 3) * 16, n0 = (warp >> 2) * 16;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float* o = sDh + (m0 + g8) * LDH + n0 + 8 * j + 2 * t4;
                *reinterpret_cast<float2*>(o) = make_float2(dhreg[j][0], dhreg[j][1]);
                *reinterpret_cast<float2*>(o + 8 * LDH) = make_float2(dhreg[j][2], dhreg[j][3]);
            }
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int r = rl + 4 * i, ag = sAgent[r];
            if (ag >= 0) atomicAdd(gr.dA + (size_t)ag * H + u, sDh[r * LDH + u]);
        }
        {   // dWz[u][z] += sum_r dh0[r][u] z[r][z]   (padding rows have dh0 = 0 and z = 0)
#pragma unroll 4
            for (int r = 0; r < ROWS; ++r) {
                const float d = sDh[r * LDH + z_u];
                awz0 = fmaf(d, sZ[r * ZMAX + z_g], awz0);
                awz1 = fmaf(d, sZ[r * ZMAX + z_g + 8], awz1);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            sDu[prow * LDU + mq + 4 * j] = dbs[j];
            ab1[j] += dbs[j];
        }
        stage_matrix(sW1, LDH, w.W1s + (size_t)g * M1 * H, M1, H);
        for (int i = threadIdx.x; i < ROWS * H; i += MGGAN_THREADS) {
            int r = i / H, k = i - r * H, ag = sAgent[r];
            sHt[r * LDH + k] = ag >= 0 ? __ldg(social + (size_t)ag * H + k) : 0.f;
        }
        __syncthreads();
        if (pag >= 0) {     // dsocial[agent][k] += sum_m W1s[m][k] dBs[row][m],  k = mq*8 .. +7
            float ds[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) ds[k] = 0.f;
#pragma unroll
            for (int m = 0; m < M1; ++m) {
                float d = sDu[prow * LDU + m];
                float4 wa = ld4(sW1 + m * LDH + mq * 8), wb = ld4(sW1 + m * LDH + mq * 8 + 4);
                ds[0] = fmaf(d, wa.x, ds[0]); ds[1] = fmaf(d, wa.y, ds[1]); ds[2] = fmaf(d, wa.z, ds[2]); ds[3] = fmaf(d, wa.w, ds[3]);
                ds[4] = fmaf(d, wb.x, ds[4]); ds[5] = fmaf(d, wb.y, ds[5]); ds[6] = fmaf(d, wb.z, ds[6]); ds[7] = fmaf(d, wb.w, ds[7]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) atomicAdd(gr.dsocial + (size_t)pag * H + mq * 8 + k, ds[k]);
        }
        {   // dW1s: rows [8 warp, 8 warp + 8) of dBs^T social, merged in shared memory
            float part[4][4];
#pragma unroll
            for (int a = 0; a < 4; ++a) { part[a][0] = 0.f; part[a][1] = 0.f; part[a][2] = 0.f; part[a][3] = 0.f; }
            tile_wgrad<8>(part, sDu + warp * 8 * LDU, LDU, e_mq * 4, sHt + warp * 8 * LDH, LDH, e_kq * 4);
            atomic_block44(sW1sAcc, H, e_mq * 4, e_kq * 4, part);
        }
    }
    __syncthreads();
    if (cur_g >= 0) flush(cur_g);
}

size_t dec_fwd_smem() {
    return sizeof(float) * (4 * H * LDH + 2 * M1 * LDH + 2 * ROWS * LDH + ROWS * LDU + ROWS * 2 + H * (ZMAX + 1));
}
size_t dec_bwd_smem() {
    return sizeof(float) * (KG * LDG + M1 * LDH + ROWS * LDG + 4 * ROWS * LDH + ROWS * LDU + ROWS * ZMAX + M1 * H + M1 * LDH + 2 * M1);
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

}  // namespace

extern "C" int mggan_decoder_fwd(int n_tiles, const int* tile_gen, const int* seq_agent, const int* seq_noise,
                                 const int* seq_out, const float* A, const float* social, const float* last_xy,
                                 const float* last_dxdy, const float* noise, int Z, const float* Wz, const float* Wx,
                                 const float* b, const float* Whh, const float* W1h, const float* W1s, const float* b1,
                                 const float* W2, const float* b2, int pred_len, int n_cols, float* out_abs,
                                 float* out_rel, float* acts, float* u1save, float* h0save, cudaStream_t stream) {
    MGGAN_REQUIRE(Z >= 1 && Z <= ZMAX, "mggan_decoder_fwd: noise_dim %d not in [1, %d]", Z, ZMAX);
    MGGAN_REQUIRE(pred_len >= 1 && n_tiles >= 0 && (n_tiles & 1) == 0, "mggan_decoder_fwd: bad pred_len / odd n_tiles");
    MGGAN_REQUIRE((acts == nullptr) == (u1save == nullptr) && (acts == nullptr) == (h0save == nullptr),
                  "mggan_decoder_fwd: save buffers must be all set or all null");
    if (n_tiles == 0) return MGGAN_OK;
    DecSeq sq{tile_gen, seq_agent, seq_noise, seq_out};
    DecWeights w{Wz, Wx, b, Whh, W1h, W1s, b1, W2, b2};
    size_t sm = dec_fwd_smem();
    cudaFuncSetAttribute(decoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int grid = n_tiles < sm_count() * 3 ? n_tiles : sm_count() * 3;
    decoder_fwd_kernel<<<grid, MGGAN_THREADS, sm, stream>>>(sq, n_tiles, A, social, last_xy, last_dxdy, noise, Z, w,
                                                            pred_len, n_cols, out_abs, out_rel, acts, u1save, h0save);
    return mggan_check_launch("decoder_fwd");
}

extern "C" int mggan_decoder_bwd(int n_tiles, const int* tile_gen, const int* seq_agent, const int* seq_noise,
                                 const int* seq_out, const float* social, const float* last_dxdy, const float* noise,
                                 int Z, const float* Wz, const float* Wx, const float* b, const float* Whh,
                                 const float* W1h, const float* W1s, const float* b1, const float* W2, const float* b2,
                                 int pred_len, int n_cols, const float* out_rel, const float* acts, const float* u1save,
                                 const float* h0save, const float* d_abs, const float* d_rel, float* dWz, float* dWx,
                                 float* db, float* dWhh, float* dW1h, float* dW1s, float* db1, float* dW2, float* db2,
                                 float* dA, float* dsocial, cudaStream_t stream) {
    MGGAN_REQUIRE(Z >= 1 && Z <= ZMAX, "mggan_decoder_bwd: noise_dim %d not in [1, %d]", Z, ZMAX);
    MGGAN_REQUIRE(acts && u1save && h0save, "mggan_decoder_bwd: forward was run without save buffers");
    MGGAN_REQUIRE((n_tiles & 1) == 0, "mggan_decoder_bwd: odd n_tiles (work lists are padded to 128-row groups)");
    if (n_tiles == 0) return MGGAN_OK;
    DecSeq sq{tile_gen, seq_agent, seq_noise, seq_out};
    DecWeights w{Wz, Wx, b, Whh, W1h, W1s, b1, W2, b2};
    DecGrads gr{dWz, dWx, db, dWhh, dW1h, dW1s, db1, dW2, db2, dA, dsocial};
    size_t sm = dec_bwd_smem();
    cudaFuncSetAttribute(decoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    int grid = n_tiles < sm_count() * 2 ? n_tiles : sm_count() * 2;
    decoder_bwd_kernel<<<grid, MGGAN_THREADS, sm, stream>>>(sq, n_tiles, social, last_dxdy, noise, Z, w, pred_len,
                                                            n_cols, out_rel, acts, u1save, h0save, d_abs, d_rel, gr);
    return mggan_check_launch("decoder_bwd");
}
