import torch, sys
sys.path.insert(0,'.')
from redtail_b200 import ops
def run(n,c,h,w,d):
    l=torch.randn(n,c,h,w,device='cuda'); r=torch.randn(n,c,h,w,device='cuda')
    try:
        cv=ops.cost_volume(l,r,d); torch.cuda.synchronize()
        ok = torch.equal(cv[:, :, :c], l[:, None].expand(-1, d, -1, -1, -1))
        print((n,c,h,w,d),'ok',ok, flush=True)
    except Exception as e:
        print((n,c,h,w,d),'FAIL',e, flush=True); sys.exit(1)
for cfg in [(1,2,8,513,4),(1,2,16,513,4),(1,2,161,513,4),(1,32,161,513,4),(1,2,161,513,48),(1,32,161,513,48)]:
    run(*cfg)
