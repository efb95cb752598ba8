"""GPU scratch tool: measured errors of the training step against tests/golden/train.npz (prints, does not assert)
and timing of a c5-shaped step (32 videos).   python scripts/probe_train.py"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_train as tt
from mucon_b200 import train
from mucon_b200.loss import mucon_loss_batch

dev = torch.device("cuda:0")
m, mods = tt._cuda_model(dev)
feats, lengths, trs = tt.golden_inputs()
print("fresh:", tt.fresh(mods, feats))
plan = m.plan(tt.TS, dev)
packed = torch.cat([f[0] for f in feats]).to(dev)
len_cat = torch.cat(lengths).to(dev).requires_grad_(True)
tr_cat = torch.cat(trs).to(dev)
seg, _ = train.forward_train_packed(m, packed, plan)
loss = mucon_loss_batch(len_cat, seg, tr_cat, tt.NS, tt.TS)
loss.backward()
print("loss", loss.item(), "want", float(tt.G["loss"]))
for k, p in m.named_parameters():
    maxabs, rms, rel = tt.grad_errors(p.grad.cpu().numpy(), k)
    print(f"{k:32s} max|err|/rms {maxabs / max(rms, 1e-30):9.2e}  rel_l2 {rel:9.2e}  rms {rms:9.2e}")

# the same against the TF32-emulating oracle (oracle.backbone._Tf32Conv), on the GPU in fp32
import torch.nn.functional as F
from oracle import backbone as obb, loss as oloss
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
got = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
for p in m.parameters():
    p.grad = None
sd = dict(m.named_parameters())
total = 0.0
for f, l, tr in zip(feats, lengths, trs):
    f = f.to(dev)
    z = obb.encode(sd, f, tt.STAGES, tt.POOL, tf32=True)
    sg = F.conv1d(F.interpolate(z.permute(0, 2, 1), f.shape[1]), sd["conv_classifier.weight"], sd["conv_classifier.bias"]).squeeze(0).permute(1, 0)
    total = total + oloss.mucon_loss(l.to(dev), sg, tr.to(dev), "box", 0.0, "flint") / len(tt.TS)
total.backward()
print("vs TF32-emulating oracle: loss", loss.item(), total.item())
worst = 0.0
for k, p in m.named_parameters():
    want = p.grad.double()
    if want.abs().max().item() == 0:
        continue
    rel = ((got[k].double() - want).norm() / want.norm()).item()
    worst = max(worst, rel)
    print(f"{k:32s} rel_l2 {rel:9.2e}  max|err|/max|want| {(got[k].double() - want).abs().max().item() / want.abs().max().item():9.2e}")
print("worst rel_l2", worst)

# ---- timing of a c5-shaped step: 32 videos, T from c2's distribution, D = 2048 -------------------------------------
from mucon_b200.temporal import MuConBackbone
import bench
T_all, _, _ = bench.make_split(0)
rng = np.random.default_rng(5)
Ts = [int(t) for t in T_all[:32]]
Ns = [int(rng.integers(2, 13)) for _ in Ts]
torch.manual_seed(0)
m2 = MuConBackbone().to(dev).train()
m2.ft.dropout_rate = 0.25
plan2 = m2.plan(Ts, dev)
f2 = torch.randn(int(sum(Ts)), 2048, device=dev).abs() * 0.5
l2 = torch.randn(int(sum(Ns)), device=dev, requires_grad=True)
t2 = torch.from_numpy(np.concatenate([rng.integers(0, 48, n) for n in Ns])).to(dev)
from mucon_b200.loss import _flint_meta
meta = _flint_meta(Ns, Ts, dev)
def step():
    for p in m2.parameters():
        p.grad = None
    seg, _ = train.forward_train_packed(m2, f2, plan2)
    loss = mucon_loss_batch(l2, seg, t2, Ns, Ts, meta=meta)
    loss.backward()
    return loss
for _ in range(3):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
frames = sum(Ts)
print(f"c5 step (32 videos, {frames} frames, dropout 0.25): {ms:.3f} ms  -> {3 * 1.014e6 * frames / ms / 1e9:.1f} TFLOP/s (3 x fwd flops)")
# the same step captured in a CUDA graph (train.TrainStep), with an SGD step inside
opt = torch.optim.SGD(m2.parameters(), lr=1e-3)
ts = train.TrainStep(m2, Ts, Ns, optimizer=opt)
ts.feats.copy_(f2); ts.transcripts.copy_(t2)
with torch.no_grad():
    ts.lengths.copy_(l2)
ts.run(); torch.cuda.synchronize()
e0.record()
for _ in range(20):
    ts.run()
e1.record(); torch.cuda.synchronize()
msg = e0.elapsed_time(e1) / 20
print(f"c5 step as a CUDA graph (+ SGD): {msg:.3f} ms -> {3 * 1.014e6 * frames / msg / 1e9:.1f} TFLOP/s, loss {ts.loss.item():.4f}")
# the reference's own way on the same GPU: its modules restated with torch (cudnn), one video at a time
sd2 = dict(m2.named_parameters())
def ref_step():
    for p in m2.parameters():
        p.grad = None
    total, o, r = 0.0, 0, 0
    for T, n in zip(Ts, Ns):
        z = obb.encode(sd2, f2[o:o + T][None], m2.ft.stages, m2.ft.pooling_layers)
        sg = F.conv1d(F.interpolate(z.permute(0, 2, 1), T), sd2["conv_classifier.weight"], sd2["conv_classifier.bias"]).squeeze(0).permute(1, 0)
        total = total + oloss.mucon_loss(l2[r:r + n], sg, t2[r:r + n], "box", 0.0, "flint") / len(Ts)
        o, r = o + T, r + n
    total.backward()
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
for _ in range(2):
    ref_step()
torch.cuda.synchronize()
e0.record()
for _ in range(3):
    ref_step()
e1.record(); torch.cuda.synchronize()
print(f"torch eager (cudnn, TF32 allowed) per-video loop, same batch: {e0.elapsed_time(e1) / 3:.3f} ms")
