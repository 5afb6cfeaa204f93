import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch, ddb200
from ddb200 import systems
A,X,Y = systems.linear_snapshots(6,1000,seed=6,restart=10**9)
x = np.concatenate([X, Y[:, -1:]], axis=1)
res = ddb200.solve_dmd(ddb200.DiscreteDataDrivenProblem(x), ddb200.B200DMD(ddb200.DMDPINV()))
print(np.round(np.real(res.K),3))
P=X@X.T; Q=Y@X.T
print("P err", np.abs(res.P-P).max(), "Q err", np.abs(res.Q-Q).max())
# direct operator call with numpy P,Q uploaded
ctx = ddb200.Context(0)
dev=torch.device("cuda:0")
Pt=torch.from_numpy(np.asfortranarray(P).T.copy()).to(dev); Qt=torch.from_numpy(np.asfortranarray(Q).T.copy()).to(dev)
Kd=torch.zeros((6,6),dtype=torch.float64,device=dev)
ctx.dmd_operator(Pt.data_ptr(), Qt.data_ptr(), 6, Kd.data_ptr())
print(np.round(Kd.T.cpu().numpy(),3))
print(np.round(A,3))
