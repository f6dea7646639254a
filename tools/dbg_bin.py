import sys, os
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "."))
import numpy as np, torch
from pnfft_b200 import api as A
N=(32,)*3; M=20000
rng=np.random.default_rng(0)
x=torch.empty((M,3),dtype=torch.float64,pin_memory=True); x.numpy()[...]=rng.uniform(-0.5,0.4999,(M,3))
fh=torch.empty(N+(2,),dtype=torch.float64,pin_memory=True); fh.numpy()[...]=rng.standard_normal(N+(2,))
f=torch.empty((M,2),dtype=torch.float64,pin_memory=True); g=torch.empty((M,3,2),dtype=torch.float64,pin_memory=True)
comm=A.create_procmesh_2d(1,1)
plan=A.Plan.init_guru(N,(64,)*3,(0.5,)*3,6,0,comm)
nd=A.Nodes(M,0); nd.set_x(x); nd.set_f(f); nd.set_grad_f(g)
for it in range(2):
    plan.set_f_hat(fh); plan.trafo(nd,3); plan.adj(nd,1)
