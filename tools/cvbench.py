"""GPU helper: times the dense (plugin-layout) cost-volume kernel on the NVSmall shape with CUDA events."""
import sys, torch
sys.path.insert(0, ".")
from redtail_b200 import ops
l = torch.randn(1, 32, 161, 513, device="cuda"); r = torch.randn(1, 32, 161, 513, device="cuda")
for _ in range(3): cv = ops.cost_volume(l, r, 48)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(20): ops.cost_volume(l, r, 48)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 20
print("cost_volume dense: %.3f ms  %.0f GB/s (algorithmic 1036.0 MB)" % (ms, 1036.0e6 / ms / 1e6), ops.last_kernel())
