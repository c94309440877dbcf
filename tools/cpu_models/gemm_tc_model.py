# numpy model of gemm_tc_kernel's slab loop, staging layout, padding, split-K and epilogue (the tcgen05.mma itself is
# modelled as "read the (rows x 32) operand planes through the canonical K-major layout and multiply": the hardware side of
# that statement is what decoder_tc.cu measured), checked against A @ B^T.
import numpy as np
KC, TROWS = 32, 128
def oper_off(r, k): return (r >> 3) * 256 + (k >> 2) * 32 + (r & 7) * 4 + (k & 3)
def tf32_hi(x): return (x.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
def run(A, sam, sak, B, sbn, sbk, M, N, K, splitk=1, colsum=False):
    npad = (N + 15) // 16 * 16
    C = np.zeros((M, N)); cs = np.zeros(M)
    rr, kk = np.meshgrid(np.arange(TROWS), np.arange(KC), indexing="ij")
    rb, kb = np.meshgrid(np.arange(npad), np.arange(KC), indexing="ij")
    for bz in range(splitk):
        k_begin, k_end = 0, K
        if splitk > 1:
            per = (K + splitk - 1) // splitk; per = (per + KC - 1) // KC * KC
            k_begin = bz * per; k_end = min(K, k_begin + per)
            if k_begin >= k_end: continue
        for bx in range((M + TROWS - 1) // TROWS):
            D = np.zeros((TROWS, npad)); csum = np.zeros(TROWS)
            for k0 in range(k_begin, k_end, KC):
                sAhi = np.full(TROWS * KC, np.nan, np.float32); sAlo = sAhi.copy()
                sBhi = np.full(npad * KC, np.nan, np.float32); sBlo = sBhi.copy()
                for tid in range(TROWS):
                    m = bx * TROWS + tid
                    for kq in range(KC // 4):
                        v = np.zeros(4, np.float32)
                        for c in range(4):
                            k = k0 + kq * 4 + c
                            if m < M and k < k_end: v[c] = A[m * sam + k * sak]
                        csum[tid] += v.sum()
                        h = tf32_hi(v)
                        for c in range(4):
                            sAhi[oper_off(tid, kq * 4 + c)] = h[c]; sAlo[oper_off(tid, kq * 4 + c)] = v[c] - h[c]
                    for n in range(tid, npad, TROWS):
                        for kq in range(KC // 4):
                            v = np.zeros(4, np.float32)
                            for c in range(4):
                                k = k0 + kq * 4 + c
                                if n < N and k < k_end: v[c] = B[n * sbn + k * sbk]
                            h = tf32_hi(v)
                            for c in range(4):
                                sBhi[oper_off(n, kq * 4 + c)] = h[c]; sBlo[oper_off(n, kq * 4 + c)] = v[c] - h[c]
                Ah, Al = sAhi[oper_off(rr, kk)].astype(np.float64), sAlo[oper_off(rr, kk)].astype(np.float64)
                Bh, Bl = sBhi[oper_off(rb, kb)].astype(np.float64), sBlo[oper_off(rb, kb)].astype(np.float64)
                D += Ah @ Bh.T + Al @ Bh.T + Ah @ Bl.T
            for tid in range(TROWS):
                m = bx * TROWS + tid
                if m >= M: continue
                C[m, :] += D[tid, :N]
                if colsum: cs[m] += csum[tid]
    return C, cs
rng = np.random.default_rng(1)
for (M, N, K, splitk) in [(300, 96, 192, 1), (37, 64, 24, 1), (129, 1, 96, 1), (65, 8, 65, 1), (96, 192, 1000, 3), (32, 64, 777, 4), (127, 8, 3, 1)]:
    A = rng.normal(size=(M, K)).astype(np.float32); B = rng.normal(size=(N, K)).astype(np.float32)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    C, _ = run(A.ravel(), K, 1, B.ravel(), K, 1, M, N, K, splitk)
    assert np.abs(C - ref).max() <= 2e-5 * np.abs(ref).max(), (M, N, K, np.abs(C - ref).max())
    At, Bt = np.ascontiguousarray(A.T), np.ascontiguousarray(B.T)
    C, cs = run(At.ravel(), 1, M, Bt.ravel(), 1, N, M, N, K, max(splitk, 2), colsum=True)
    assert np.abs(C - ref).max() <= 2e-5 * np.abs(ref).max() and np.allclose(cs, A.astype(np.float64).sum(1), atol=1e-4)
    print((M, N, K, splitk), "max rel err", np.abs(C - ref).max() / np.abs(ref).max())
print("gemm_tc_kernel model ok")
