import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import parity
from pointcloud_bridge_b200 import synthetic
from pointcloud_bridge_b200.engine import Trainer
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg
g = parity.load("models.npz"); DEV='cuda:0'
x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"])).to(DEV)
lab = torch.from_numpy(g["labels"].astype(np.int64)).to(DEV)
def run(mode, lr=1e-3):
    torch.manual_seed(7)
    net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV).train(); net.drop1.eval()
    tr = Trainer(net, amp=False, graph=mode, capturable=True, lr=lr)
    torch.manual_seed(11)
    return [float(tr.step(x9, labels=lab).item()) for _ in range(6)]
print("eager A", run(False)); print("eager B", run(False)); print("graph  ", run(True))
print("lr=0: eager", run(False, 0.0)); print("lr=0: graph", run(True, 0.0))
