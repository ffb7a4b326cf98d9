import sys, numpy as np, torch
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import parity
from pointcloud_bridge_b200 import synthetic
from pointcloud_bridge_b200.partsize import pointnet2_sem_seg_msg as msg
torch.backends.cuda.matmul.allow_tf32=False
g = parity.load("models.npz")
DEV='cuda:0'
x9 = torch.from_numpy(synthetic.sem_seg_input(g["xyz"], g["rgb"])).to(DEV)
lab = torch.from_numpy(g["labels"].astype(np.int64)).to(DEV)
def rel(a, ref):
    a=a.detach().float().cpu().numpy(); return float(np.abs(a-ref).max()/(np.abs(ref).max()+1e-12))
net = parity.seeded_fill_(msg.get_model(5), 2).to(DEV); net.train(); net.drop1.eval()
torch.manual_seed(4242)
y,_ = net(x9)
loss = torch.nn.functional.nll_loss(y.reshape(-1,5), lab.reshape(-1)); loss.backward()
print("loss", loss.item(), float(g["msg_train_loss"]), "logp", rel(y, g["msg_train_logp"]))
for name,p in (("g_sa1", net.sa1.conv_blocks[0][0].weight.grad), ("g_fp1", net.fp1.mlp_convs[0].weight.grad), ("g_conv2", net.conv2.weight.grad), ("rm_sa1", net.sa1.bn_blocks[0][0].running_mean), ("rv_sa1", net.sa1.bn_blocks[0][0].running_var)):
    print(name, rel(p, g[f"msg_train_{name}"]), np.abs(g[f"msg_train_{name}"]).max())
# same in float64 on GPU for the MLP part? compare double-run determinism
net2 = parity.seeded_fill_(msg.get_model(5), 2).to(DEV); net2.train(); net2.drop1.eval()
torch.manual_seed(4242)
y2,_ = net2(x9); l2 = torch.nn.functional.nll_loss(y2.reshape(-1,5), lab.reshape(-1)); l2.backward()
print("run-to-run g_sa1", rel(net2.sa1.conv_blocks[0][0].weight.grad, net.sa1.conv_blocks[0][0].weight.grad.cpu().numpy()))
