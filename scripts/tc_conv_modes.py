import sys, numpy as np, torch, subprocess
sys.path.insert(0, "/root/repo")
mode, K = sys.argv[1], int(sys.argv[2])
import dspeed_b200.processors as P
from oracle import oracle as O
rows, L = 150, 4096
rng = np.random.default_rng(K)
x = rng.normal(0, 4, (rows, L)).astype(np.float32); k = rng.standard_normal(K).astype(np.float32)
P.TC_CONV_MIN_TAPS = 1
p = {"v": L - K + 1, "f": L + K - 1, "s": L}[mode]
out = torch.empty((rows, p), dtype=torch.float32, device="cuda")
P.convolve_wf(torch.from_numpy(x).cuda(), torch.from_numpy(k).cuda(), np.int8(ord(mode)), out); torch.cuda.synchronize()
ref = O.convolve_wf(x, k, mode)
print(mode, K, "err", np.abs(out.cpu().numpy() - ref).max() / np.abs(ref).max())
