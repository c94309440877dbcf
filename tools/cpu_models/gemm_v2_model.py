# numpy model of gemm_kernel_v2's index math (thread-vectorised, barrier phases explicit), checked against A @ B^T
import numpy as np
BK, BN, T = 16, 64, 256
def run(Aflat, sam, sak, Bflat, sbn, sbk, M, N, K, TM, splitk=1, colsum=False):
    BM2 = 16*TM; NA = BM2*BK//T; NB = BN*BK//T
    C = np.zeros((M, N)); cs = np.zeros(M)
    tid = np.arange(T); tm = tid >> 4; tn = tid & 15
    a_kfast, b_kfast = sak == 1, sbk == 1
    for bz in range(splitk):
        k_begin, k_end = 0, K
        if splitk > 1:
            per = (K + splitk - 1)//splitk; per = (per + BK - 1)//BK*BK
            k_begin = bz*per; k_end = min(K, k_begin+per)
            if k_begin >= k_end: continue
        for bx in range((M+BM2-1)//BM2):
            for by in range((N+BN-1)//BN):
                m0, n0 = bx*BM2, by*BN
                acc = np.zeros((T, TM, 4)); csum = np.zeros(T)
                def fetch(k0):
                    ra = np.zeros((NA, T)); rb = np.zeros((NB, T))
                    for it in range(NA):
                        idx = tid + it*T
                        mm = idx//BK if a_kfast else idx % BM2; kk = idx % BK if a_kfast else idx//BM2
                        m = m0+mm; k = k0+kk; ok = (m < M) & (k < k_end)
                        ra[it, ok] = Aflat[(m*sam + k*sak)[ok]]
                    for it in range(NB):
                        idx = tid + it*T
                        nn = idx//BK if b_kfast else idx % BN; kk = idx % BK if b_kfast else idx//BN
                        n = n0+nn; k = k0+kk; ok = (n < N) & (k < k_end)
                        rb[it, ok] = Bflat[(n*sbn + k*sbk)[ok]]
                    return ra, rb
                ra, rb = fetch(k_begin)
                for k0 in range(k_begin, k_end, BK):
                    As = np.full((BK, BM2+4), np.nan); Bs = np.full((BK, BN+4), np.nan)
                    for it in range(NA):
                        idx = tid + it*T
                        mm = idx//BK if a_kfast else idx % BM2; kk = idx % BK if a_kfast else idx//BM2
                        As[kk, mm] = ra[it]
                    for it in range(NB):
                        idx = tid + it*T
                        nn = idx//BK if b_kfast else idx % BN; kk = idx % BK if b_kfast else idx//BN
                        Bs[kk, nn] = rb[it]
                    if k0 + BK < k_end: ra, rb = fetch(k0+BK)
                    for kk in range(BK):
                        a = np.stack([As[kk, tm*TM + i] for i in range(TM)], 1)
                        b = np.stack([Bs[kk, tn*4 + j] for j in range(4)], 1)
                        acc += a[:, :, None]*b[:, None, :]
                    if colsum and by == 0:
                        sel = tid < BM2
                        csum[sel] += As[:, tid[sel]].sum(0)
                for t in range(T):
                    for i in range(TM):
                        m = m0 + tm[t]*TM + i
                        if m >= M: continue
                        for j in range(4):
                            n = n0 + tn[t]*4 + j
                            if n >= N: continue
                            C[m, n] += acc[t, i, j]
                if colsum and by == 0:
                    for t in range(min(T, BM2)):
                        if m0 + t < M: cs[m0+t] += csum[t]
    return C, cs
rng = np.random.default_rng(0)
for (M, N, K, TM, splitk) in [(300, 96, 192, 8, 1), (37, 64, 24, 4, 1), (129, 1, 96, 8, 1), (65, 8, 65, 8, 1), (96, 192, 1000, 8, 3), (32, 64, 777, 4, 4), (127, 8, 3, 8, 1)]:
    A = rng.normal(size=(M, K)); B = rng.normal(size=(N, K))
    # forward-like: A k-fast, B k-fast
    C, _ = run(A.ravel(), K, 1, B.ravel(), K, 1, M, N, K, TM, splitk)
    assert np.allclose(C, A @ B.T), ("kfast", M, N, K)
    # weight-gradient-like: A(m,k) = dY[k*O + m] (sam=1, sak=M), B(n,k) = X[k*Kx + n] (sbn=1, sbk=N), colsum
    At = np.ascontiguousarray(A.T); Bt = np.ascontiguousarray(B.T)
    C, cs = run(At.ravel(), 1, M, Bt.ravel(), 1, N, M, N, K, TM, max(splitk, 2), colsum=True)
    assert np.allclose(C, A @ B.T) and np.allclose(cs, A.sum(1)), ("strided", M, N, K)
    # input-gradient-like: A k-fast, B(n,k) = W[k*Kw + n]
    C, _ = run(A.ravel(), K, 1, Bt.ravel(), 1, N, M, N, K, TM, 1)
    assert np.allclose(C, A @ B.T)
print("gemm_kernel_v2 index model ok")
