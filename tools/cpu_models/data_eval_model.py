# line-by-line numpy transliteration of the three data_eval.cu kernels, checked against the reference-frozen fixtures
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


f32 = np.float32
z = np.load(ROOT + '/tests/golden/scene_crop.npz')
imgs = [z[f'crop/image{i}'] for i in range(int(z['crop/n_images']))]
atlas = np.concatenate([im.reshape(-1) for im in imgs]); off = np.cumsum([0] + [im.size for im in imgs])[:-1]
wh = [(im.shape[1], im.shape[0]) for im in imgs]
ids = z['crop/image_id']; xy = z['crop/last_xy']; scal = z['crop/scaling_small']
scale = {i: f32(1.0 / scal[list(ids).index(i)]) for i in range(len(imgs))}
CROP, PIX, MARG = 33, 1089, 16
for a in range(len(ids)):
    im = int(ids[a]); w, h = wh[im]; s = scale[im]
    x0 = int(f32(xy[a, 0]) * s) - MARG; y0 = int(f32(xy[a, 1]) * s) - MARG       # C cast truncates toward zero like int()
    out = np.empty(4 * PIX, f32)
    for i in range(4 * PIX):
        c = i // PIX; p = i - c * PIX
        if c == 3:
            v = f32(1.0) if p == MARG * CROP + MARG else f32(0.0)
        else:
            yy = p // CROP; y = y0 + yy; x = x0 + (p - yy * CROP); u = 0
            if 0 <= x < w and 0 <= y < h: u = int(atlas[off[im] + (y * w + x) * 3 + c])
            v = f32(f32(u) * f32(2.0 / 256.0) + f32(-1.0))
        out[i] = v
    assert np.array_equal(out.reshape(4, 33, 33), z['crop/features'][a]), a
print("crop kernel transliteration == reference crops:", len(ids))

e = np.load(ROOT + '/tests/golden/evaluation.npz')
man, tests, mask = e['inside/manifold'], e['inside/tests'], e['inside/mask']
T = man.shape[1]; r = float(e['meta/radius']); radius = np.linspace(r / T, r, T)
pool = np.concatenate([man, tests]); m = len(man)
res = []
for i in range(len(tests)):
    tp = pool[m + i]; all_t = True
    for t in range(T):
        any_ = False
        for j in range(m):
            mp = pool[j, t]
            dx = f32(mp[0] - tp[t, 0]); dy = f32(mp[1] - tp[t, 1])
            d = np.sqrt(f32(f32(dx * dx) + f32(dy * dy)), dtype=f32)
            any_ |= float(d) < radius[t]
        if not any_: all_t = False; break
    res.append(all_t)
assert np.array_equal(np.array(res), mask)
print("tube kernel transliteration == reference mask:", int(mask.sum()), "of", mask.size)

# min ADE/FDE
import torch
preds = e['pred/abs']; gt = e['batch/gt_xy']
ok = ~np.isnan(gt).any(-1).any(0)
sse = e['batch/seq_start_end']; offs = np.concatenate([[0], np.cumsum(ok)])
se = [(s - (s - offs[s]) , 0) for s, _ in sse]
start_end = [(int(offs[s]), int(offs[e_])) for s, e_ in sse]
p = preds[:, :, ok]; g = gt[:, ok]; Tn, K, n, _ = p.shape
ade = np.zeros((len(start_end), K)); fde = np.zeros_like(ade); mode = np.zeros_like(ade, dtype=int)
for sc, (a0, a1) in enumerate(start_end):
    sa = np.zeros(K); sf = np.zeros(K); sm = np.zeros(K, int)
    for i in range(a0, a1):
        best = np.inf
        for s in range(K):
            sm_ = f32(0)
            for t in range(Tn):
                dx = p[t, s, i, 0] - g[t, i, 0]; dy = p[t, s, i, 1] - g[t, i, 1]
                ee = np.sqrt(f32(dx * dx + dy * dy)); sm_ = f32(sm_ + ee)
            sa[s] += sm_; sf[s] += ee; best = min(best, ee)
            if best < 3.0: sm[s] += 1
    ade[sc] = np.minimum.accumulate(sa); fde[sc] = np.minimum.accumulate(sf); mode[sc] = sm
for k in range(1, K + 1):
    assert abs(ade[:, k-1].sum() / (Tn * n) - float(e[f'ade_fde/ADE k={k}'])) < 1e-5
    assert abs(fde[:, k-1].sum() / n - float(e[f'ade_fde/FDE k={k}'])) < 1e-5
    assert abs(mode[:, k-1].sum() / n - float(e[f'ade_fde/Mode k={k}'])) < 1e-9
print("min ADE/FDE kernel transliteration == reference metrics")
