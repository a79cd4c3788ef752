"""Host logic of `texpose_b200.model.nerf_pretrain.Graph` on the CPU: the kernel wrappers it calls (`camera.get_center_and_ray`,
`ops.gather_rows`, `ops.sample_depth`, `NeRF.forward_samples`, `NeRF.composite`) are swapped for the oracle's restatements -- test
infrastructure standing in for the CUDA library, which has no CPU path -- so that what is exercised is the drop-in's own
orchestration: ray-subset draw, gathers, depth-range packing, slicing of a frame, the `ray_idx=None` call form, the three losses and
their autograd.  Everything is held to the fixture of the REAL reference (tests/golden/pretrain.npz: model/nerf_pretrain.py
`forward(mode='train')` + `compute_loss` + backward, `render_by_slices(mode='val')`), whose CPU generator draws this run reproduces."""
import pytest
import torch

from oracle import texpose_oracle as O
from texpose_b200 import camera, ops
from texpose_b200.config import AttrDict, env_opt
from texpose_b200.layers.nerf import NeRF
from texpose_b200.model import nerf_pretrain


@pytest.fixture
def oracle_kernels(monkeypatch):
    def center_and_ray(opt, pose, intr=None, H=None, W=None, ray_idx=None):
        c, r = O.get_center_and_ray(pose, intr, opt.H, opt.W)
        return (c, r) if ray_idx is None else (O.gather_rays(c, ray_idx), O.gather_rays(r, ray_idx))

    def sample_depth(zn, zf, N, rand=None, stratified=True, seed=None):
        return O.sample_depth(zn, zf, N, rand if stratified else None)

    def forward_samples(self, opt, center, ray, depth_samples, mode=None):
        pts = O.points_from_depth(center, ray, depth_samples)
        unit = torch.nn.functional.normalize(ray, dim=-1)[..., None, :].expand_as(pts)
        layers = lambda ml: [(l.weight, l.bias) for l in ml]
        return O.nerf_plain_forward(pts, unit, layers(self.mlp_feat), layers(self.mlp_rgb))

    monkeypatch.setattr(camera, "get_center_and_ray", center_and_ray)
    monkeypatch.setattr(ops, "gather_rows", lambda src, idx: O.gather_rays(src.float(), idx))
    monkeypatch.setattr(ops, "sample_depth", sample_depth)
    monkeypatch.setattr(NeRF, "forward_samples", forward_samples)
    monkeypatch.setattr(NeRF, "composite", staticmethod(lambda opt, ray, rgb, dens, depth: O.composite_plain(ray, rgb, dens, depth)))


def _graph(g):
    opt = env_opt(H=int(g.H), W=int(g.W), sample_intvs=int(g.N))
    opt.nerf.rand_rays = g.ray_idx.numel()
    opt.loss_weight = AttrDict(render=0, mask=-1, depth=-1)
    opt.data.erode_mask_loss = False
    torch.manual_seed(0)
    return opt, nerf_pretrain.Graph(opt)


def close(a, b, tol=2e-6):
    assert a.shape == b.shape and (a.double() - b.double()).abs().max().item() <= tol, (a.double() - b.double()).abs().max().item()


def test_train_forward_loss_and_gradients_match_the_reference_fixture(golden, oracle_kernels):
    g = golden("pretrain")
    opt, graph = _graph(g)
    B = g.ray_idx.shape[0]
    var = AttrDict(idx=torch.arange(B), pose=g.pose, pose_init=g.pose, intr=g.intr, z_near=g.z_near, z_far=g.z_far, image=g.image,
                   obj_mask=g.obj_mask, depth_gt=g.depth_gt)
    torch.manual_seed(21)                         # the seed the reference ran with: same randperm subset, same jitter
    var = graph.forward(opt, var, mode="train")
    assert torch.equal(var.ray_idx, g.ray_idx)
    close(var.rgb, g.o_rgb); close(var.depth, g.o_depth, 5e-6); close(var.opacity, g.o_opacity)
    loss = graph.compute_loss(opt, var, mode="train")
    assert set(loss.keys()) == {"mask", "depth", "render"}
    for k in loss:
        assert abs(loss[k].item() - float(g["l_" + k])) <= 1e-6, k
    total = sum(10 ** float(opt.loss_weight[k]) * loss[k] for k in loss)
    assert abs(total.item() - float(g.l_total)) <= 1e-6
    total.backward()
    for n, p in graph.nerf.named_parameters():
        got = p.grad if p.grad.numel() <= 2048 else p.grad[:, ::8][::8]
        close(got, g["g/" + n])


def test_frame_slices_and_the_no_ray_list_call_match_the_reference_fixture(golden, oracle_kernels):
    g = golden("pretrain")
    opt, graph = _graph(g)
    opt.nerf.sample_stratified = False
    dr = (g.z_near[:1, :, None], g.z_far[:1, :, None])
    with torch.no_grad():
        whole = graph.render_by_slices(opt, g.pose[:1], intr=g.intr[:1], depth_range=dr, object_mask=g.obj_mask[:1], mode="val")
        opt.b200 = AttrDict(slice_rays=5000)      # 12 288 rays -> 5000 + 5000 + 2288
        parts = graph.render_by_slices(opt, g.pose[:1], intr=g.intr[:1], depth_range=dr, mode="val")       # no mask, as :417-420 calls it
        direct = graph.render(opt, g.pose[:1], intr=g.intr[:1], depth_range=dr)                            # no ray list, as :421-422
    for k in ("rgb", "depth", "opacity"):
        close(whole[k], g["v_" + k], 5e-6 if k == "depth" else 2e-6)
        close(parts[k], g["v_" + k], 5e-6 if k == "depth" else 2e-6)
        close(direct[k], g["v_" + k], 5e-6 if k == "depth" else 2e-6)
    # val / eval modes take the ground-truth pose, the env engine always (model/nerf_pretrain_env.py:484-485)
    var = AttrDict(idx=torch.arange(1), pose=g.pose[:1], pose_init=g.pose[1:2], intr=g.intr[:1], z_near=g.z_near[:1], z_far=g.z_far[:1],
                   obj_mask=g.obj_mask[:1])
    with torch.no_grad():
        out = graph.forward(opt, var, mode="val")
    close(out.rgb, g.v_rgb)
