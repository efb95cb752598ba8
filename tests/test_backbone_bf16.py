"""GPU parity of the bf16 WaveNet layer kernel (csrc/backbone_bf16.cuh) and of the bf16 backbone path.

Layer kernel: against a torch fp32 evaluation of the SAME bf16-rounded operands (x, weights, and the bf16-rounded
intermediate relu(conv)), so the only differences are the accumulation order inside the tensor core and one final
rounding to bf16: a bf16 half-ulp relative (2^-8) plus a small absolute term.
Whole backbone: against the frozen outputs of the unmodified reference WaveNetBlock (tests/golden/backbone.npz),
per precision, with the measured error printed (run with -s)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.backbone_util import CASES, G, POOL, STAGES, case_inputs, state_dict_of

pytestmark = pytest.mark.gpu


def rms(x):
    return float(np.sqrt(np.mean(np.square(x))))


def layer_ref(x16, wd16, bd, w116, b1, Ts, dil, pool, relu_final):
    """x16 [sum T, 128] bf16 (cpu) -> list of per-video fp32 outputs (before the final bf16 rounding)."""
    outs, o = [], 0
    wd = wd16.float().view(3, 128, 128).permute(1, 2, 0).contiguous()   # [Cout, Cin, k]
    w1 = w116.float().view(128, 128, 1)
    for T in Ts:
        x = x16[o:o + T].float().t()[None]                               # [1, 128, T]
        o += T
        y = F.relu(F.conv1d(x, wd, bd, dilation=dil, padding=dil))
        y = y.to(x16.dtype).float()
        z = F.conv1d(y, w1, b1) + x
        if relu_final:
            z = F.relu(z)
        if pool:
            z = F.max_pool1d(z, 2) if T >= 2 else z[:, :, :0]
        outs.append(z[0].t().contiguous())
    return outs


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("dil", [1, 2, 4, 8, 16, 32, 64, 128, 1024])
def test_bf16_layer_kernel(cuda_device, dil, dt):
    from mucon_b200.temporal import BackbonePlan, wavenet_layer_bf16_rows
    g = torch.Generator().manual_seed(11 + dil)
    Ts = [700, 333, 64, 1999, 16, 128, 129, 127, 256, 257, 1024, 17, 2048, 300, 1, 2, 3]
    plan = BackbonePlan(Ts, 1, cuda_device)
    x16 = torch.randn(sum(Ts), 128, generator=g).to(dt)
    wd16 = (torch.randn(3 * 128, 128, generator=g) / 20).to(dt)
    w116 = (torch.randn(128, 128, generator=g) / 11).to(dt)
    bd, b1 = torch.randn(128, generator=g), torch.randn(128, generator=g)
    dev = cuda_device
    for pool, relu_final, out_f32 in [(False, False, False), (True, False, False), (False, True, True), (True, True, False)]:
        got = wavenet_layer_bf16_rows(x16.to(dev), wd16.to(dev), bd.to(dev), w116.to(dev), b1.to(dev), plan, 0, dil,
                                      pool, relu_final, out_f32=out_f32)
        torch.cuda.synchronize()
        assert got.dtype == (torch.float32 if out_f32 else dt)
        got = got.float().cpu()
        ref = layer_ref(x16, wd16, bd, w116, b1, Ts, dil, pool, relu_final)
        off = plan.off_host[1 if pool else 0]
        for v, T in enumerate(Ts):
            a, b = got[off[v]:off[v + 1]], ref[v]
            assert a.shape == b.shape, (v, T, a.shape, b.shape)
            if a.numel() == 0:
                continue
            # the bf16-rounded intermediate can flip by one ulp when the accumulation order differs: absolute slack
            tol = 2e-2 + (0 if out_f32 else 2.0 ** -8) * b.abs() if dt == torch.bfloat16 else 3e-3 + 2.0 ** -11 * b.abs()
            bad = ((a - b).abs() > tol)
            assert not bad.any(), (dil, pool, relu_final, out_f32, v, T, (a - b).abs().max().item(),
                                   bad.nonzero()[:4].tolist())


@pytest.mark.parametrize("dt", [torch.bfloat16, torch.float16])
def test_last_conv_on_the_layer_pipeline(cuda_device, dt):
    """conv1x1_bf16_rows (the layer kernel without its skip connection, identity centre tap, dead side taps): a plain
    1x1 conv of non-negative 16-bit rows -> fp32, against torch fp32 on the same rounded operands.  The identity GEMM
    and the 16-bit round trip of relu(x) are exact, so only the accumulation order differs: 1e-4 relative to the row
    scale.  With the skip connection (residual=True) the same call adds x back."""
    from mucon_b200.temporal import DEAD_DILATION, BackbonePlan, conv1x1_bf16_rows, wavenet_layer_bf16_rows
    g = torch.Generator().manual_seed(5)
    Ts = [700, 333, 64, 1999, 16, 128, 129, 127, 1, 2, 3, 140, 19]
    plan = BackbonePlan(Ts, 1, cuda_device)
    x16 = torch.randn(sum(Ts), 128, generator=g).relu().to(dt)
    w16 = (torch.randn(128, 128, generator=g) / 11).to(dt)          # [Cout, Cin]
    b = torch.randn(128, generator=g)
    ident = torch.zeros(3 * 128, 128)
    ident[128:256] = torch.eye(128)
    dev = cuda_device
    got = conv1x1_bf16_rows(x16.to(dev), ident.to(dt).to(dev), w16.to(dev), b.numpy(), plan, 0)
    torch.cuda.synchronize()
    assert got.dtype == torch.float32 and got.shape == (sum(Ts), 128)
    want = x16.double() @ w16.double().t() + b.double()
    assert (got.cpu().double() - want).abs().max().item() <= 1e-4 * want.abs().max().item()
    skip = wavenet_layer_bf16_rows(x16.to(dev), ident.to(dt).to(dev), np.zeros(128, np.float32), w16.to(dev), b.numpy(),
                                   plan, 0, DEAD_DILATION, False, False, out_f32=True, residual=True)
    assert (skip.cpu().double() - (want + x16.double())).abs().max().item() <= 1e-4 * want.abs().max().item()


@pytest.mark.parametrize("precision", ["fp16", "bf16", "tf32"])
@pytest.mark.parametrize("i", range(len(CASES)))
def test_reference_golden_end_to_end_precisions(cuda_device, i, precision):
    """Frozen outputs of the unmodified reference modules; prints the measured error so the stated bars can be
    audited in the test log."""
    from mucon_b200.temporal import MuConBackbone
    (T, D, H, C), (ft, gn, cls), feats, fresh = case_inputs(i)
    if not fresh:
        pytest.skip("torch RNG stream differs from the one the fixture was minted with")
    if H != 128:
        pytest.skip("tensor-core paths are built for 128 channels")
    m = MuConBackbone(input_feature_size=D, num_classes=C, hidden_size=H).eval()
    m.load_state_dict(state_dict_of(ft, gn, cls))
    m = m.to(cuda_device)
    plan = m.plan([T])
    z = m.encode_packed(feats[0].to(cuda_device).contiguous(), plan, precision=precision)
    want_z = G[f"c{i}_z"]
    err = np.abs(z.cpu().numpy() - want_z)
    logp = m.logprobs_packed(z, plan).cpu().numpy()
    want = G[f"c{i}_logp"]
    got = logp if T <= 800 else logp[::7]
    errl = np.abs(got - want)
    print(f"\n[{precision}] case {i} T={T} D={D}: z max|err|/RMS={err.max() / rms(want_z):.4f} rms(err)/RMS="
          f"{rms(err) / rms(want_z):.5f}; logp max|err|/RMS={errl.max() / rms(want):.4f} rms(err)/RMS="
          f"{rms(errl) / rms(want):.5f} argmax agree={np.mean(np.argmax(got, 1) == np.argmax(want, 1)):.4f}")
    # stated bars (max |err| and RMS error, both relative to the RMS of the reference tensor): tf32: max <= 2e-2
    # (measured <= 1.1e-2); fp16: max <= 3e-2 = SURVEY.md 8c's bar (measured <= 2.2e-2; the maximum is one element
    # of a small-variance GroupNorm group and moves between 1.5e-2 and 2.2e-2 with the order in which the layer
    # kernel adds the bias, while the RMS error stays at 1.1e-3 to 1.3e-3); RMS error <= 2e-3 for both; bf16: rounding the residual stream to an 8-bit
    # mantissa at every layer gives a 1e-2 RMS error and single elements (small-variance GroupNorm groups) up
    # to 0.1 * RMS -- outside SURVEY.md 8c's 3e-2, which is why fp16 is the default 16-bit type
    max_bar, rms_bar, agree = {"fp16": (3e-2, 2e-3, 0.995), "tf32": (2e-2, 2e-3, 0.995), "bf16": (1.5e-1, 1.5e-2, 0.97)}[precision]
    assert rms(err) <= rms_bar * rms(want_z) and rms(errl) <= rms_bar * rms(want)
    assert err.max() <= max_bar * rms(want_z)
    assert errl.max() <= max_bar * rms(want)
    assert np.mean(np.argmax(got, 1) == np.argmax(want, 1)) >= agree


@pytest.mark.parametrize("precision", ["fp16", "bf16"])
def test_bf16_backbone_ragged_batch_matches_single_videos(cuda_device, precision):
    """A packed ragged batch gives the same activations as the same videos one at a time (tile walk, padding,
    pooling floors and the slab / separate-tap modes of every layer)."""
    from mucon_b200.temporal import MuConBackbone
    torch.manual_seed(5)
    m = MuConBackbone(input_feature_size=64, num_classes=20).eval().to(cuda_device)
    Ts = [700, 333, 64, 1999, 16, 128, 129, 127, 4100, 257, 1024, 17, 2048, 300]
    feats = [torch.randn(t, 64).abs().to(cuda_device) for t in Ts]
    plan = m.plan(Ts)
    z = m.encode_packed(torch.cat(feats), plan, precision=precision)
    zo = plan.off_host[-1]
    for v, t in enumerate(Ts):
        zv = m.encode_packed(feats[v], m.plan([t]), precision=precision)
        assert torch.equal(z[zo[v]:zo[v + 1]], zv), (v, t, (z[zo[v]:zo[v + 1]] - zv).abs().max().item())


def test_full_inference_pooled_equals_expanded(cuda_device):
    """Backbone -> alignment, two ways: (a) classifier + log-softmax expanded to [sum T, C] and aligned from that
    array (the drop-in's materialising path), (b) log-softmax at the pooled resolution and the fused alignment kernel
    reading that table through the nearest-neighbour index.  Scores, segments and labels must be bit-identical."""
    from mucon_b200.length_model import poisson_params
    from mucon_b200.temporal import MuConBackbone
    from mucon_b200.viterbi import AlignPlan, ViterbiEngine
    from tests import synth
    torch.manual_seed(1)
    rng = np.random.default_rng(1)
    net = MuConBackbone(input_feature_size=64, num_classes=48).eval().to(cuda_device)
    Ts = [int(t) for t in rng.integers(300, 3000, 24)] + [16 * 30 + 7, 30, 31]
    trs, means = [], []
    for t in Ts:
        K = t // 30
        n = int(rng.integers(max(1, -(-K // 66)), min(12, K) + 1))
        tr = rng.integers(0, 48, n).tolist()
        trs.append(tr)
        means.append(synth.class_means(rng.dirichlet(np.ones(n)).astype(np.float32), tr, 48, t))
    feats = torch.randn(sum(Ts), 64, device=cuda_device).abs()
    bplan = net.plan(Ts)
    z = net.encode_packed(feats, bplan)
    eng = ViterbiEngine(cuda_device)
    params = np.stack([poisson_params(m) for m in means])
    res = []
    for pooled in (False, True):
        plan = AlignPlan(Ts, [[tr] for tr in trs], 48, device=cuda_device, labels="best", len_params=params)
        if pooled:
            lsm, zoff = net.logprobs_pooled_packed(z, bplan)
            eng.run(plan, lsm, seg0_f32=True, z_off=zoff)
        else:
            eng.run(plan, net.logprobs_packed(z, bplan), seg0_f32=True, mode="fused")
        torch.cuda.synchronize()
        res.append(eng.fetch(plan))
    for k in ("score", "status", "seg_blocks", "labels", "final_j"):
        assert np.array_equal(res[0][k], res[1][k]), k
    assert (res[0]["status"] == 0).all()


@pytest.mark.parametrize("ncls", [20, 48, 64])
def test_fused_tail_equals_separate_kernels(cuda_device, ncls):
    """mucon_tail_logprobs (GroupNorm statistics + one launch for GroupNorm / ReLU / classifier / log_softmax) against
    the separate GroupNorm, classifier and log-softmax kernels on a ragged batch."""
    from mucon_b200.temporal import MuConBackbone
    torch.manual_seed(2)
    net = MuConBackbone(input_feature_size=64, num_classes=ncls).eval().to(cuda_device)
    with torch.no_grad():
        net.ft_last_gn.weight.uniform_(0.5, 1.5)
        net.ft_last_gn.bias.uniform_(-0.5, 0.5)
    Ts = [700, 333, 64, 1999, 16, 128, 129, 127, 4100, 257, 1024, 17, 2048, 300]
    feats = torch.cat([torch.randn(t, 64).abs() for t in Ts]).to(cuda_device)
    plan = net.plan(Ts)
    table, off, z = net.infer_pooled_packed(feats, plan, want_z=True)
    z2 = net.encode_packed(feats, plan)
    table2, off2 = net.logprobs_pooled_packed(z2, plan)
    assert torch.equal(off, off2)
    assert torch.allclose(z, z2, rtol=1e-5, atol=1e-5), (z - z2).abs().max().item()
    assert torch.allclose(table, table2, rtol=1e-5, atol=2e-5), (table - table2).abs().max().item()
    assert torch.allclose(torch.logsumexp(table, 1), torch.zeros_like(table[:, 0]), atol=1e-5)


@pytest.mark.gpu
def test_pipelined_inference_is_bit_identical(cuda_device):
    """infer_pooled_pipelined (projection and layer kernels of consecutive video chunks side by side on disjoint SMs,
    mucon_set_sm_limit) returns exactly infer_pooled_packed's table"""
    import numpy as np
    from mucon_b200.temporal import MuConBackbone
    torch.manual_seed(5)
    m = MuConBackbone(input_feature_size=256).eval().to(cuda_device)
    T = np.array([700, 333, 64, 1999, 16, 128, 129, 1024, 17, 300, 2500, 900])
    plan = m.plan(T, cuda_device)
    feats = torch.randn(int(T.sum()), 256, device=cuda_device).abs() * 0.5
    want, off = m.infer_pooled_packed(feats, plan)
    got, off2 = m.infer_pooled_pipelined(feats, T, n_chunks=3, proj_sms=80)
    torch.cuda.synchronize()
    assert torch.equal(got, want) and torch.equal(off, off2)
