import sys, numpy as np, torch
sys.path.insert(0, "/root/repo")
import dspeed_b200.processors as P
from oracle import oracle as O
for rows, L, K in [(256, 8192, 256), (200, 8192, 1000), (128, 8192, 4096), (256, 8192, 2048)]:
    rng = np.random.default_rng(K + rows)
    x = (rng.normal(0, 4, (rows, L)) + 2000 * (np.arange(L)[None, :] > rng.integers(L // 4, 3 * L // 4, (rows, 1)))).astype(np.float32)
    k = rng.standard_normal(K).astype(np.float32)
    ref64 = np.stack([np.convolve(x[r].astype(np.float64), k.astype(np.float64), "valid") for r in range(min(rows, 32))])
    res = {}
    for tc in (1, 0):
        P.TC_CONV_MIN_TAPS = 1 if tc else 0
        xd = torch.from_numpy(x).cuda(); out = torch.empty((rows, L - K + 1), dtype=torch.float32, device="cuda")
        P.convolve_wf(xd, torch.from_numpy(k).cuda(), np.int8(ord("v")), out); torch.cuda.synchronize()
        res[tc] = out.cpu().numpy()[:32]
    ora = O.convolve_wf(x[:32], k, "v")
    sc = np.abs(ref64).max()
    print(K, "scale %.3g" % sc, "tc %.2e direct %.2e oracle(f32) %.2e  (max abs err / scale vs float64)" % (
        np.abs(res[1] - ref64).max() / sc, np.abs(res[0] - ref64).max() / sc, np.abs(ora - ref64).max() / sc))
