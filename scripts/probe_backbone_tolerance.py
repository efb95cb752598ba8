import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from oracle import backbone as ob
from tests.backbone_util import STAGES, POOL
from mucon_b200.temporal import MuConBackbone
def rms(x): return float(np.sqrt(np.mean(np.square(x))))
dev=torch.device('cuda:0')
torch.manual_seed(3)
m = MuConBackbone(input_feature_size=48, num_classes=20).eval()
with torch.no_grad():
    m.ft_last_gn.weight.uniform_(0.5, 1.5); m.ft_last_gn.bias.uniform_(-0.5, 0.5)
sd = {k: v.clone() for k, v in m.state_dict().items()}
Ts = [700, 333, 64, 1999, 16, 128, 129, 127, 256, 257, 1024, 17, 2048, 300]
feats = [torch.randn(1, t, 48).abs() for t in Ts]
mc = m.to(dev); plan = mc.plan(Ts)
packed = torch.cat([f[0] for f in feats]).to(dev)
for name, kw in dict(tf32_unfused=dict(tensor_cores=True, fused_layers=False, precision="tf32"), tf32_fused=dict(precision="tf32"), fp16=dict(precision="fp16"), bf16=dict(precision="bf16")).items():
    z = mc.encode_packed(packed, plan, **kw); logp = mc.logprobs_packed(z, plan)
    zo, lo = plan.off_host[-1], plan.off_host[0]
    out=[]
    for v,t in enumerate(Ts):
        with torch.no_grad():
            rz = ob.encode(sd, feats[v], STAGES, POOL); rl = ob.logprobs(sd, rz, t)
        gz = z[zo[v]:zo[v+1]].cpu(); gl = logp[lo[v]:lo[v+1]].cpu()
        out.append((t, round((gz-rz[0]).abs().max().item()/rms(rz.numpy()),4), round((gl-rl).abs().max().item()/rms(rl.numpy()),4)))
    print(name, out)
