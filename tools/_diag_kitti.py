import numpy as np, torch, sys
sys.path.insert(0,".")
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
from oracle import c_oracle as co
def rel(a,b): return float(np.abs(a-b).max()/np.abs(b).max())
rs=np.random.RandomState(312)
shape=(2,48,96,312)
x1=rs.standard_normal(shape).astype(np.float32); x2=rs.standard_normal(shape).astype(np.float32)
t1,t2=torch.from_numpy(x1).cuda(),torch.from_numpy(x2).cuda()
ref=co.level_forward(x1,x2,None,8,1,8,1,1,0,0.1)
for v in (0,1,2,5):
    out=ops.warp_corr_forward(t1,t2,None,8,1,8,1,1,1,0,0.1,variant=v)
    d=np.abs(out.cpu().numpy()-ref)
    bad=np.argwhere(d>1e-4*np.abs(ref).max())
    print("variant",v,"fwd rel",rel(out.cpu().numpy(),ref), "nbad",len(bad),"planes", np.unique(bad[:,1])[:12], "rows",np.unique(bad[:,2])[:12],"cols", np.unique(bad[:,3])[:12])
g=rs.standard_normal(ref.shape).astype(np.float32)
out=ops.warp_corr_forward(t1,t2,None,8,1,8,1,1,1,0,0.1,variant=5)
g1,g2,_=ops.warp_corr_backward(t1,t2,None,out,torch.from_numpy(g).cuda(),8,1,8,1,1,1,0,0.1)
r1,r2,_=co.level_backward(x1,x2,None,g,8,1,8,1,1,0,0.1)
print("bwd", rel(g1.cpu().numpy(),r1), rel(g2.cpu().numpy(),r2))
d=np.abs(g1.cpu().numpy()-r1); bad=np.argwhere(d>1e-4*np.abs(r1).max()); print("g1 nbad",len(bad), "rows",np.unique(bad[:,2])[:12],"cols",np.unique(bad[:,3])[:12])
d=np.abs(g2.cpu().numpy()-r2); bad=np.argwhere(d>1e-4*np.abs(r2).max()); print("g2 nbad",len(bad), "rows",np.unique(bad[:,2])[:12],"cols",np.unique(bad[:,3])[:12])
