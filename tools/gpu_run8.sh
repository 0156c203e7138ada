mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_trailnet.py -x -q 2>&1 | tail -30) > gpurun_out/t_tn.log
(timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8) > gpurun_out/t_all.log
cat > /tmp/tnprof.py <<'PY'
import sys, numpy as np, torch, time
sys.path.insert(0, ".")
from redtail_b200 import CaffeNet
TN = "tests/golden/trailnet/"
B = int(sys.argv[1])
net = CaffeNet(TN + "TrailNet_SResNet-18.prototxt", TN + "TrailNet_SResNet-18.caffemodel", "out", max_batch=B)
x = torch.rand(B, 3, 180, 320).cuda() * 255
rows = net.profile(x)
for n, ms in rows: print("  %-60s %.3f" % (n[:60], ms))
print("layers", len(rows), "sum ms", sum(m for _, m in rows))
for _ in range(3): net(x)
torch.cuda.synchronize(); t0 = time.time()
for _ in range(10): net(x)
torch.cuda.synchronize(); dt = (time.time() - t0) / 10
print("batch %d: %.3f ms/step, %.0f images/s" % (B, dt * 1e3, B / dt))
PY
(timeout 300 python /tmp/tnprof.py 256 2>&1 | tail -75) > gpurun_out/tn_prof.log
cat gpurun_out/t_tn.log gpurun_out/t_all.log; tail -5 gpurun_out/tn_prof.log
