"""fp32-parity mode on the tensor cores (csrc/mlp_tc_split.cu: split-bf16 operands, three tcgen05.mma passes per K step)
against the CPU oracle and against the SIMT fp32 kernels.  Tolerance: the north-star's 1e-4 max-abs of the fp32 path on
every rendered output; per-sample head outputs 1e-4 (bounded ones) / 1e-4 relative to the range (raw densities)."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.layers.nerf_static_transient_light import NeRF
from tests.test_gpu_tc import _c1_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4
NAMES = ["rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient", "prob",
         "uncert", "alpha_static", "alpha_transient"]


def _module(engine, **arch):
    opt = adapt_gan_opt(device=DEV)
    for k, v in arch.items():
        opt.arch[k] = v
    opt.b200 = AttrDict(mlp="fp32", fp32_engine=engine)
    torch.manual_seed(0)
    return opt, NeRF(opt).to(DEV)


def _oracle_render(m, center, ray, depth, lt, ll, skip=(4,)):
    cpu_layers = lambda ml: [(l.weight.detach().cpu(), l.bias.detach().cpu()) for l in ml]
    pts = O.points_from_depth(center, ray, depth)
    unit = torch.nn.functional.normalize(ray, dim=-1)[..., None, :].expand_as(pts)
    ref_s = O.nerf_stl_forward(pts, unit, lt, ll, cpu_layers(m.mlp_feat), cpu_layers(m.mlp_rgb), cpu_layers(m.mlp_trans), skip=skip)
    return ref_s, O.composite_stl(ray, *ref_s[:2], depth, ref_s[2], 0.05)


def _randomize_biases(m, seed=12):
    gen = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for lin in list(m.mlp_feat) + list(m.mlp_rgb) + list(m.mlp_trans):
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=gen).mul_(0.3).to(DEV))


def test_c1_render_split_kernel_vs_oracle():
    """BASELINE configs[0] shape: 1024 AABB rays x 64 samples, all eleven rendered outputs within 1e-4 of the oracle."""
    center, ray, depth = _c1_inputs()
    lt, ll = synth.latents(1)
    opt, m = _module("auto")
    _randomize_biases(m)
    ref_s, ref = _oracle_render(m, center, ray, depth, lt, ll)
    with torch.no_grad():
        got_s = m.forward_samples(opt, center.to(DEV), ray.to(DEV), depth.to(DEV), lt.to(DEV), ll.to(DEV), mode="val")
        got = m.composite(opt, ray.to(DEV), *got_s[:2], depth.to(DEV), got_s[2])
    errs = {k: (a.cpu() - b).abs().max().item() for k, a, b in zip(NAMES, got, ref)}
    print("split-bf16 render max-abs errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    for k in NAMES:
        assert errs[k] <= TOL, (k, errs[k])
    assert (got_s[0].cpu() - ref_s[0]).abs().max() <= TOL
    assert (got_s[1].cpu() - ref_s[1]).abs().max() <= TOL * max(1.0, ref_s[1].abs().max().item())
    assert (got_s[2].cpu() - ref_s[2]).abs().max() <= TOL * max(1.0, ref_s[2].abs().max().item())


def test_split_kernel_is_the_default_for_rendering_and_simt_for_training():
    from texpose_b200 import _C
    center, ray, depth = [t.to(DEV) for t in _c1_inputs(R=64, N=32)]
    lt, ll = [t.to(DEV) for t in synth.latents(1)]
    opt, m = _module("auto")
    _C.launch_counts.clear()
    with torch.no_grad():
        m.forward_samples(opt, center, ray, depth, lt, ll, mode="val")
    assert _C.launch_counts.get("tp_tc32_forward") == 1 and "tp_linear_forward" not in _C.launch_counts
    _C.launch_counts.clear()
    out = m.forward_samples(opt, center, ray, depth, lt, ll, mode="train")      # gradients wanted: SIMT forward + saved activations
    assert "tp_tc32_forward" not in _C.launch_counts and _C.launch_counts.get("tp_linear_forward", 0) > 0
    out[0].sum().backward()
    opt_s, _ = _module("simt")
    _C.launch_counts.clear()
    with torch.no_grad():
        m.forward_samples(opt_s, center, ray, depth, lt, ll, mode="val")
    assert "tp_tc32_forward" not in _C.launch_counts


def test_split_kernel_matches_simt_on_ragged_batch():
    """Two images, 37 rays x 24 samples: S = 1776 is not a multiple of the tile, rays and images straddle tiles and warps."""
    B, R, N = 2, 37, 24
    g = torch.Generator().manual_seed(4)
    center = (torch.randn(B, R, 3, generator=g) * 0.02 + torch.tensor([0.3, 0.2, -0.8])).to(DEV)
    ray = (torch.randn(B, R, 3, generator=g) * 0.1 + torch.tensor([0.0, 0.0, 1.0])).to(DEV)
    depth = ((torch.rand(B, R, N, 1, generator=g) + torch.arange(N)[None, None, :, None]) / N * 1.2 + 0.2).to(DEV)
    lt, ll = [t.to(DEV) for t in synth.latents(B)]
    opt_tc, m = _module("auto")
    opt_simt, _ = _module("simt")
    _randomize_biases(m)
    with torch.no_grad():
        a = m.forward_samples(opt_tc, center, ray, depth, lt, ll, mode="val")
        b = m.forward_samples(opt_simt, center, ray, depth, lt, ll, mode="val")
        a2 = m.forward_samples(opt_tc, center, ray, depth, lt, ll, mode="val")
    assert a[0].shape == (B, R, N, 3, 2) and a[1].shape == (B, R, N, 2) and a[2].shape == (B, R, N, 1)
    for x, y in zip(a, b):
        assert (x - y).abs().max() <= TOL * max(1.0, y.abs().max().item())
    for x, y in zip(a, a2):
        assert torch.equal(x, y)            # deterministic
    with torch.no_grad():                   # weights changed in place -> the split image is rebuilt (param._version key)
        m.mlp_rgb[3].bias.add_(0.5)
        c = m.forward_samples(opt_tc, center, ray, depth, lt, ll, mode="val")
        d = m.forward_samples(opt_simt, center, ray, depth, lt, ll, mode="val")
    assert (c[0][..., 0] - a[0][..., 0]).abs().max() > 0.05
    assert (c[0] - d[0]).abs().max() <= TOL


@pytest.mark.parametrize("layers_feat,skip,layers_rgb,layers_trans", [
    ([None, 256, 256, 256, 256, 256, 256], [2], [None, 256, 256, 3], [None, 256, 5]),            # 6-layer trunk, skip at 2
    ([None, 256, 256, 256, 256, 256, 256, 256, 256, 256, 256], [3, 6], [None, 256, 3], [None, 256, 256, 256, 256, 5]),
])
def test_stage_list_follows_the_architecture(layers_feat, skip, layers_rgb, layers_trans):
    """The stage list is built from opt.arch: other trunk depths / skip sets / head depths run on the same kernel."""
    center, ray, depth = _c1_inputs(R=96, N=32)
    lt, ll = synth.latents(1)
    opt, m = _module("auto", layers_feat=layers_feat, skip=skip, layers_rgb=layers_rgb, layers_trans=layers_trans)
    _randomize_biases(m)
    from texpose_b200 import _C
    ref_s, ref = _oracle_render(m, center, ray, depth, lt, ll, skip=tuple(skip))
    _C.launch_counts.clear()
    with torch.no_grad():
        got_s = m.forward_samples(opt, center.to(DEV), ray.to(DEV), depth.to(DEV), lt.to(DEV), ll.to(DEV), mode="val")
        got = m.composite(opt, ray.to(DEV), *got_s[:2], depth.to(DEV), got_s[2])
    assert _C.launch_counts.get("tp_tc32_forward") == 1
    for k, a, b in zip(NAMES, got, ref):
        assert (a.cpu() - b).abs().max() <= TOL, k


def test_static_only_stops_after_the_rgb_head():
    center, ray, depth = [t.to(DEV) for t in _c1_inputs(R=128, N=32)]
    lt, ll = [t.to(DEV) for t in synth.latents(1)]
    opt, m = _module("auto")
    opt.b200.static_only = True
    with torch.no_grad():
        full = m.forward_samples(opt, center, ray, depth, lt, ll, mode="val")
        stat = m.forward_samples(opt, center, ray, depth, lt, ll, mode="eval")
    assert torch.equal(stat[0][..., 0], full[0][..., 0]) and torch.equal(stat[1][..., 0], full[1][..., 0])
    assert stat[0][..., 1].abs().max() == 0 and stat[1][..., 1].abs().max() == 0 and stat[2].abs().max() == 0


def test_stage_list_is_validated():
    import ctypes
    from texpose_b200 import _C, ops
    lib = _C.load()
    dummy = torch.zeros(4096, device=DEV)
    img = torch.zeros(17 * 16384, dtype=torch.uint8, device=DEV)

    def run(rows, n_slots):
        st = torch.tensor(rows, dtype=torch.int32)
        return lib.tp_tc32_forward(ops._p(dummy), ops._p(dummy), ops._p(dummy), 128, 32, 128, ops._p(img), n_slots, ops._p(st),
                                   len(rows), ops._p(dummy), ops._p(dummy), ops._p(dummy), ops._p(dummy), ops._p(dummy),
                                   ops._p(dummy), ops._p(torch.zeros(lib.tp_tc32_scratch_bytes(), dtype=torch.uint8, device=DEV)),
                                   lib.tp_tc32_scratch_bytes(), 0, None, 0, -1, ctypes.c_void_p(0))

    ok = [[0, 1, 0, 0, 0, 8, -1], [16, 0, 2, 0, 256, 1, -1]]
    assert run(ok, 2) == 0
    assert run(ok, 3) == -2                                                     # slot count does not match the list
    assert run([[0, 1, 0, 0, 0, 8, -1], [16, 0, 2, 0, 256, 0, -1]], 2) == -1            # output stage does not wait for the drain
    assert run([[0, 1, 0, 0, 0, 8, -1], [16, 0, 0, 0, 0, 2, -1], [16, 0, 2, 0, 256, 1, -1]], 18) == -1      # reload without a parked feature
    assert run([[0, 1, 0, 0, 0, 0, -1], [16, 0, 2, 0, 256, 1, -1]], 2) == -1            # nobody marks the last reader of the encoding
    torch.cuda.synchronize()


def test_plain_model_renders_on_the_split_kernel():
    """layers/nerf.py (nerf_lm_env.yaml: 128-wide rgb head, no latents) in the fp32 mode, no gradient: the split kernel with the
    head zero-padded to 256 columns, within 1e-4 of the oracle; with gradients wanted the SIMT kernels stay."""
    from texpose_b200 import _C
    from texpose_b200.config import env_opt
    from texpose_b200.layers.nerf import NeRF as PlainNeRF
    center, ray, depth = _c1_inputs(R=200, N=48)
    opt = env_opt(device=DEV)
    opt.b200 = AttrDict(mlp="fp32")
    torch.manual_seed(0)
    m = PlainNeRF(opt).to(DEV)
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for lin in list(m.mlp_feat) + list(m.mlp_rgb):
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=gen).mul_(0.3).to(DEV))
    cpu_layers = lambda ml: [(l.weight.detach().cpu(), l.bias.detach().cpu()) for l in ml]
    pts = O.points_from_depth(center, ray, depth)
    unit = torch.nn.functional.normalize(ray, dim=-1)[..., None, :].expand_as(pts)
    ref_rgb, ref_sigma = O.nerf_plain_forward(pts, unit, cpu_layers(m.mlp_feat), cpu_layers(m.mlp_rgb))
    _C.launch_counts.clear()
    with torch.no_grad():
        rgb, sigma = m.forward_samples(opt, center.to(DEV), ray.to(DEV), depth.to(DEV), mode="val")
    assert _C.launch_counts.get("tp_tc32_forward") == 1 and "tp_linear_forward" not in _C.launch_counts
    assert rgb.shape == ref_rgb.shape and sigma.shape == ref_sigma.shape
    assert (rgb.cpu() - ref_rgb).abs().max() <= TOL
    assert (sigma.cpu() - ref_sigma).abs().max() <= TOL * max(1.0, ref_sigma.abs().max().item())
    _C.launch_counts.clear()
    rgb_g, _ = m.forward_samples(opt, center.to(DEV), ray.to(DEV), depth.to(DEV), mode="train")
    assert "tp_tc32_forward" not in _C.launch_counts
    assert (rgb_g - rgb).abs().max() <= TOL


def test_single_pass_bf16_covers_other_architectures():
    """opt.b200.mlp = 'bf16' / 'auto' on an architecture the lock-step kernel is not specialised for: the staged kernel in its
    single-pass bf16 mode (<= 1e-2), not the SIMT kernels."""
    from texpose_b200 import _C
    center, ray, depth = _c1_inputs(R=128, N=32)
    lt, ll = synth.latents(1)
    arch = dict(layers_feat=[None] + [256] * 6, skip=[3], layers_rgb=[None, 256, 256, 3], layers_trans=[None, 256, 256, 5])
    opt, m = _module("auto", **arch)
    _randomize_biases(m)
    ref_s, ref = _oracle_render(m, center, ray, depth, lt, ll, skip=(3,))
    for mode in ("bf16", "auto"):
        opt.b200 = AttrDict(mlp=mode)
        _C.launch_counts.clear()
        with torch.no_grad():
            got_s = m.forward_samples(opt, center.to(DEV), ray.to(DEV), depth.to(DEV), lt.to(DEV), ll.to(DEV), mode="val")
            got = m.composite(opt, ray.to(DEV), *got_s[:2], depth.to(DEV), got_s[2])
        assert _C.launch_counts.get("tp_tc32_forward") == 1 and "tp_linear_forward" not in _C.launch_counts
        errs = {k: (a.cpu() - b).abs().max().item() for k, a, b in zip(NAMES, got, ref)}
        for k in ("rgb", "rgb_static", "rgb_transient", "depth", "opacity", "opacity_static", "opacity_transient"):
            assert 1e-6 < errs[k] <= 1e-2 or errs[k] <= 1e-2, (k, errs[k])
        assert errs["uncert"] <= 1.5e-2


def test_c2_rays_split_kernel_vs_oracle():
    """BASELINE configs[1] sampling: 1024 AABB-bounded rays of the 480 x 640 frame at 128 samples per ray (one ray per tile),
    all eleven rendered outputs within 1e-4 of the oracle."""
    center, ray, depth = _c1_inputs(R=1024, N=128)
    lt, ll = synth.latents(1)
    opt, m = _module("auto")
    ref_s, ref = _oracle_render(m, center, ray, depth, lt, ll)
    with torch.no_grad():
        got_s = m.forward_samples(opt, center.to(DEV), ray.to(DEV), depth.to(DEV), lt.to(DEV), ll.to(DEV), mode="val")
        got = m.composite(opt, ray.to(DEV), *got_s[:2], depth.to(DEV), got_s[2])
    for k, a, b in zip(NAMES, got, ref):
        assert (a.cpu() - b).abs().max() <= TOL, (k, float((a.cpu() - b).abs().max()))


def test_split_mode_reports_activations_outside_the_fp16_range():
    from texpose_b200 import mlp_tc32
    center, ray, depth = [t.to(DEV) for t in _c1_inputs(R=64, N=32)]
    lt, ll = [t.to(DEV) for t in synth.latents(1)]
    opt, m = _module("auto")
    with torch.no_grad():
        m.forward_samples(opt, center, ray, depth, lt, ll, mode="val")
    mlp_tc32.check_range(torch.device(DEV))                 # a sane network: nothing to report
    with torch.no_grad():
        m.mlp_feat[2].weight.mul_(3e4)                      # hidden activations of ~1e5
        m.forward_samples(opt, center, ray, depth, lt, ll, mode="val")
    with pytest.raises(FloatingPointError):
        mlp_tc32.check_range(torch.device(DEV))
    mlp_tc32.check_range(torch.device(DEV))                 # the flag is cleared by the failed check
