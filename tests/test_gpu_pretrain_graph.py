"""`model.nerf_pretrain.Graph` (reference model/nerf_pretrain.py:449-466, :513-536, :588-660, :707-728): the pre-training
engine's render calls land on the same kernels as the adaptation engine's.  Checked against the CPU oracle on the same
seeded rays and the same jitter draw: fp32 mode <= 1e-4, bf16 mode <= 1e-2; slices concatenate to the single-launch frame
bit for bit; the training-mode forward + loss + backward runs the tensor-core step."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import _C, synth
from texpose_b200.config import AttrDict, env_opt
from texpose_b200.model.nerf_pretrain import Graph

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, W, N = 96, 128, 32


@pytest.fixture(autouse=True)
def _reference_bit_constants():
    """K^-1 / pose^-1 with the CPU reference's bits (see tests/test_gpu_render.py): end-to-end parity is checked on identical rays."""
    from texpose_b200 import camera
    camera.HOST_MATRICES = True
    yield
    camera.HOST_MATRICES = False


def _scene(B=1):
    pose = synth.poses(list(range(B)))
    intr = synth.intrinsics(B).clone()
    intr[:, :2] *= 0.2                                      # the LineMOD camera at 96 x 128
    c, r = O.get_center_and_ray(pose, intr, H, W)
    lo, hi = synth.padded_aabb()
    tn, tf, v = O.aabb_ray_intersection(lo, hi, c, r)
    z_near = torch.where(v, tn, torch.full_like(tn, 7.0))   # rays that miss the box keep a fixed range
    z_far = torch.where(v, tf, torch.full_like(tf, 9.0))
    return pose, intr, c, r, z_near, z_far, v


def _graph(mlp, stratified=True, seed=0, **b200):
    opt = env_opt(H=H, W=W, sample_intvs=N, device=DEV)
    opt.nerf.sample_stratified = stratified
    opt.b200 = AttrDict(mlp=mlp, **b200)
    torch.manual_seed(seed)
    g = Graph(opt).to(DEV)
    gen = torch.Generator().manual_seed(7)
    with torch.no_grad():
        for lin in list(g.nerf.mlp_feat) + list(g.nerf.mlp_rgb):
            lin.bias.copy_(torch.randn(lin.bias.shape, generator=gen).mul_(0.2).to(DEV))
    cpu = Graph(env_opt(H=H, W=W, sample_intvs=N))
    cpu.load_state_dict(g.state_dict())
    return opt, g, cpu


def _oracle_render(cpu, c, r, zn, zf, rand):
    depth = O.sample_depth(zn, zf, N, rand)
    pts = O.points_from_depth(c, r, depth)
    unit = torch.nn.functional.normalize(r, dim=-1)[..., None, :].expand_as(pts)
    fl = [(l.weight, l.bias) for l in cpu.nerf.mlp_feat]
    rl = [(l.weight, l.bias) for l in cpu.nerf.mlp_rgb]
    rgb_s, sig = O.nerf_plain_forward(pts, unit, fl, rl)
    return O.composite_plain(r, rgb_s, sig, depth)


@pytest.mark.parametrize("mlp,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_render_sampled_rays_vs_oracle(mlp, tol):
    B, R = 2, 300
    pose, intr, c, r, z_near, z_far, valid = _scene(B)
    opt, g, cpu = _graph(mlp)
    gen = torch.Generator().manual_seed(5)
    ray_idx = torch.randperm(H * W, generator=gen)[:R].repeat(B, 1)      # object and background rays alike
    torch.manual_seed(11)
    rand = torch.rand(B, R, N, 1, device=DEV).cpu()         # the draw render() makes after the same seed
    torch.manual_seed(11)
    _C.launch_counts.clear()
    with torch.no_grad():
        ret = g.render(opt, pose.to(DEV), intr=intr.to(DEV), ray_idx=ray_idx.to(DEV),
                       depth_range=(z_near[..., None].to(DEV), z_far[..., None].to(DEV)), mode="val")
    assert "tp_linear_forward" not in _C.launch_counts      # both precisions run on the tensor-core kernels
    ref = _oracle_render(cpu, O.gather_rays(c, ray_idx), O.gather_rays(r, ray_idx),
                         O.gather_rays(z_near[..., None], ray_idx)[..., 0], O.gather_rays(z_far[..., None], ray_idx)[..., 0], rand)
    errs = {k: (ret[k].cpu() - v).abs().max().item() for k, v in zip(("rgb", "depth", "opacity"), ref)}
    print(f"pretrain Graph.render ({mlp}) max-abs errors:", {k: f"{v:.2e}" for k, v in errs.items()})
    assert set(ret.keys()) == {"rgb", "depth", "opacity"}
    assert ret.rgb.shape == (B, R, 3) and ret.depth.shape == (B, R, 1) and ret.opacity.shape == (B, R, 1)
    for k, e in errs.items():
        assert e <= tol, (k, e)


def test_render_by_slices_is_the_single_launch_frame():
    pose, intr, c, r, z_near, z_far, _ = _scene(1)
    opt, g, cpu = _graph("bf16", stratified=False)
    args = dict(intr=intr.to(DEV), depth_range=(z_near[..., None].to(DEV), z_far[..., None].to(DEV)),
                object_mask=torch.ones(1, H, W, device=DEV), mode="val")
    with torch.no_grad():
        whole = g.render_by_slices(opt, pose.to(DEV), **args)
        opt.b200.slice_rays = 5000                          # 12288 rays -> 3 ragged slices
        parts = g.render_by_slices(opt, pose.to(DEV), **args)
    for k in ("rgb", "depth", "opacity"):
        assert whole[k].shape[1] == H * W
        assert torch.equal(whole[k], parts[k]), k
    ref = _oracle_render(cpu, c, r, z_near, z_far, None)
    for k, v in zip(("rgb", "opacity"), (ref[0], ref[2])):
        assert (whole[k].cpu() - v).abs().max() <= 1e-2, k


def test_forward_train_step_runs_on_tensor_cores_and_matches_the_oracle_loss():
    B = 2
    pose, intr, c, r, z_near, z_far, valid = _scene(B)
    opt, g, cpu = _graph("bf16")
    opt.nerf.rand_rays = 512
    opt.loss_weight = AttrDict(render=0, mask=-1, depth=-1)
    opt.data.erode_mask_loss = False
    gen = torch.Generator().manual_seed(3)
    var = AttrDict(idx=torch.arange(B), pose=pose.to(DEV), pose_init=pose.to(DEV), intr=intr.to(DEV),
                   z_near=z_near.to(DEV), z_far=z_far.to(DEV), image=torch.rand(B, 3, H, W, generator=gen).to(DEV),
                   obj_mask=valid.view(B, H, W).float().to(DEV),
                   depth_gt=(7.5 + torch.rand(B, H, W, generator=gen)).to(DEV))
    torch.manual_seed(21)
    _C.launch_counts.clear()
    var = g.forward(opt, var, mode="train")
    loss = g.compute_loss(opt, var, mode="train")
    total = sum(10 ** float(opt.loss_weight[k]) * loss[k] for k in loss)
    total.backward()
    assert _C.launch_counts.get("tp_tc32_forward") == 1 and _C.launch_counts.get("tp_tc_chain_backward") == 1
    assert "tp_linear_forward" not in _C.launch_counts
    R = opt.nerf.rand_rays // B
    assert var.ray_idx.shape == (B, R) and torch.equal(var.ray_idx[0], var.ray_idx[1])
    # ---- oracle on the rays and jitter the forward drew
    torch.manual_seed(21)
    ray_idx = torch.randperm(H * W, device=DEV)[:R].repeat(B, 1).cpu()
    assert torch.equal(ray_idx, var.ray_idx.cpu())
    rand = torch.rand(B, R, N, 1, device=DEV).cpu()
    gr = lambda t: O.gather_rays(t, ray_idx)
    rgb, depth, opacity, _ = _oracle_render(cpu, gr(c), gr(r), gr(z_near[..., None])[..., 0], gr(z_far[..., None])[..., 0], rand)
    image = gr(var.image.cpu().view(B, 3, H * W).permute(0, 2, 1))
    mask = gr(valid.view(B, H * W, 1).float())
    dgt = gr(var.depth_gt.cpu().view(B, H * W, 1))
    want = AttrDict(mask=((mask - opacity) ** 2).mean(),
                    depth=((1 - torch.minimum(depth, dgt) / (torch.maximum(depth, dgt) + 1e-5)) * mask).sum() / (mask.sum() + 1e-5),
                    render=(mask * (image - rgb) ** 2).sum() / (mask.sum() + 1e-5))
    for k in ("mask", "depth", "render"):
        assert abs(loss[k].item() - want[k].item()) <= 1e-2, (k, loss[k].item(), want[k].item())
    sum(10 ** float(opt.loss_weight[k]) * want[k] for k in want).backward()
    worst = 0.0
    for (n, a), b in zip(g.nerf.named_parameters(), cpu.nerf.parameters()):
        assert a.grad is not None, n
        worst = max(worst, (a.grad.cpu() - b.grad).abs().max().item())
    print(f"pretrain Graph train step: worst gradient max-abs error {worst:.2e}")
    assert worst <= 1e-2


@pytest.mark.parametrize("mlp,tol", [("fp32", 1e-4), ("bf16", 1e-2)])
def test_train_step_and_val_frame_match_the_real_reference(golden, monkeypatch, mlp, tol):
    """Fixture `pretrain` = model/nerf_pretrain.py Graph of the unmodified reference: forward(mode='train') + compute_loss +
    autograd and render_by_slices(mode='val').  The CPU generator's ray subset and jitter are injected (randperm / rand draw
    differently on the device); everything else goes through the drop-in's public calls."""
    g = golden("pretrain")
    B, R = g.ray_idx.shape
    opt = env_opt(H=int(g.H), W=int(g.W), sample_intvs=int(g.N), device=DEV)
    opt.nerf.rand_rays = B * R
    opt.loss_weight = AttrDict(render=0, mask=-1, depth=-1)
    opt.data.erode_mask_loss = False
    opt.b200 = AttrDict(mlp=mlp)
    torch.manual_seed(0)
    gr = Graph(opt).to(DEV)                                  # tf_init weights of seed 0 == the fixture's (checksums pinned on CPU)
    var = AttrDict(idx=torch.arange(B), pose=g.pose.to(DEV), pose_init=g.pose.to(DEV), intr=g.intr.to(DEV),
                   z_near=g.z_near.to(DEV), z_far=g.z_far.to(DEV), image=g.image.to(DEV), obj_mask=g.obj_mask.to(DEV),
                   depth_gt=g.depth_gt.to(DEV))
    real_randperm, real_rand = torch.randperm, torch.rand
    perm = torch.cat([g.ray_idx[0], torch.zeros(int(g.H) * int(g.W) - R, dtype=torch.long)]).to(DEV)
    monkeypatch.setattr(torch, "randperm", lambda n, **kw: perm if n == int(g.H) * int(g.W) else real_randperm(n, **kw))
    monkeypatch.setattr(torch, "rand", lambda *s, **kw: g.rand.to(DEV) if tuple(s) == tuple(g.rand.shape) else real_rand(*s, **kw))
    var = gr.forward(opt, var, mode="train")
    loss = gr.compute_loss(opt, var, mode="train")
    monkeypatch.undo()
    assert torch.equal(var.ray_idx.cpu(), g.ray_idx)
    sum(10 ** float(opt.loss_weight[k]) * loss[k] for k in loss).backward()
    errs = {k: (var[k].cpu() - g["o_" + k]).abs().max().item() for k in ("rgb", "depth", "opacity")}
    errs.update({"l_" + k: abs(loss[k].item() - float(g["l_" + k])) for k in ("render", "mask", "depth")})
    worst = 0.0
    for n, p in gr.nerf.named_parameters():
        got = p.grad if p.grad.numel() <= 2048 else p.grad[:, ::8][::8]
        worst = max(worst, (got.cpu() - g["g/" + n]).abs().max().item())
    errs["grad"] = worst
    with torch.no_grad():
        opt.nerf.sample_stratified = False
        val = gr.render_by_slices(opt, g.pose[:1].to(DEV), intr=g.intr[:1].to(DEV),
                                  depth_range=(g.z_near[:1, :, None].to(DEV), g.z_far[:1, :, None].to(DEV)),
                                  object_mask=g.obj_mask[:1].to(DEV), mode="val")
    for k in ("rgb", "depth", "opacity"):
        errs["val_" + k] = (val[k].cpu() - g["v_" + k]).abs().max().item()
    print(f"pretrain Graph vs the real reference ({mlp}):", {k: f"{v:.2e}" for k, v in errs.items()})
    for k, e in errs.items():
        assert e <= tol, (k, e)
