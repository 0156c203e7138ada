"""GPU helper for ncu: NVSmall 1025x321 engine, N inference steps on the bench's synthetic pair (default 3)."""
import sys, torch
sys.path.insert(0, ".")
import bench
from redtail_b200 import StereoEngine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
eng = StereoEngine("nvsmall", bench.H, bench.W, bench.WEIGHTS, max_batch=1)
l, r = bench.synthetic_pairs(1)
dl, dr = torch.from_numpy(l).cuda(), torch.from_numpy(r).cuda()
out = torch.empty((1, bench.H, bench.W), dtype=torch.float32, device="cuda")
for _ in range(n):
    eng(dl, dr, out=out)
torch.cuda.synchronize()
print("ok", float(out.mean()))
