"""GPU helper: TrailNet S-ResNet-18 at batch B -- per-step times and images/s."""
import sys, numpy as np, torch, time
sys.path.insert(0, ".")
from redtail_b200 import CaffeNet
from tests.util import trailnet_model_files
PROTO, MODEL = trailnet_model_files()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
net = CaffeNet(PROTO, MODEL, "out", max_batch=B)
x = torch.rand(B, 3, 180, 320).cuda() * 255
rows = net.profile(x)
for n, ms in rows: print("  %-60s %.3f" % (n[:60], ms))
print("layers", len(rows), "sum ms", sum(m for _, m in rows))
for _ in range(3): net(x)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(10): net(x)
torch.cuda.synchronize(); dt = (time.time() - t0) / 10
print("batch %d: %.3f ms/step, %.0f images/s" % (B, dt * 1e3, B / dt))
