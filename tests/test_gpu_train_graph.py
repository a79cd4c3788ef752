"""`train_graph.GraphedStep`: the whole texture-learner step -- Graph.render(mode='train') (model/nerf_adapt_st_gan.py:547-631),
compute_loss + summarize_loss (:712-763, model/base.py:145-157), backward, the engine's Adam step (:62-69,117-126) and the weight
re-pack -- captured once as a CUDA graph and replayed with new inputs: the replayed training run must equal the eager run of the same
steps bit for bit (every kernel of the path is deterministic), i.e. nothing of the step depends on host state frozen at capture."""
import pytest
import torch

from texpose_b200 import _C, compute_box, synth
from texpose_b200.config import AttrDict, adapt_gan_opt
from texpose_b200.model.base import summarize_loss
from texpose_b200.model.nerf_adapt_st_gan import Graph
from texpose_b200.train_graph import GraphedStep

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B, P, N, HW = 4, 8, 32, 64


def _setup():
    opt = adapt_gan_opt(H=HW, W=HW, sample_intvs=N, device=DEV)
    opt.batch_size = B
    opt.nerf.sample_stratified = False            # no jitter: the two runs need no common random stream
    opt.b200 = AttrDict(mlp="bf16", rng="torch")
    torch.manual_seed(0)
    g = Graph(opt, n_train_images=B).to(DEV).train()
    optim = torch.optim.Adam([dict(params=g.nerf.parameters(), lr=1.e-3)], capturable=True)
    optim.add_param_group(dict(params=g.latent_vars_light.parameters(), lr=1.e-3))
    optim.add_param_group(dict(params=g.latent_vars_trans.parameters(), lr=1.e-3))
    pose = synth.poses(list(range(B))).to(DEV)
    K = torch.tensor([[286.2, 0, 32 - 286.2 * 0.3 / 8], [0, 286.8, 32 + 286.8 * 0.2 / 8], [0, 0, 1]])
    intr = K.repeat(B, 1, 1).to(DEV)
    lo, hi = [t.to(DEV) for t in synth.padded_aabb()]
    zn, zf = compute_box.box_range(pose, intr, lo, hi, HW, HW, *synth.BG_RANGE)
    static = dict(coords=torch.zeros(B, P, P, 2, device=DEV), image=torch.zeros(B, 3, HW, HW, device=DEV),
                  mask=torch.zeros(B, HW, HW, device=DEV))
    idx = torch.arange(B, device=DEV)

    def fn():
        ret = g.render(opt, pose, intr=intr, ray_idx=static["coords"], depth_range=(zn[:, :, None], zf[:, :, None]),
                       sample_idx=idx, mode="train")
        var = AttrDict(idx=idx, image=static["image"], obj_mask=static["mask"], ray_idx=static["coords"])
        var.update(ret)
        loss = g.compute_loss(opt, var, mode="train")
        total = summarize_loss(opt, var, loss)["all"]
        total.backward()
        return total

    return g, optim, static, fn


def _batch(i):
    gen = torch.Generator().manual_seed(100 + i)
    return dict(coords=synth.patch_coords(B, P, seed=50 + i)[0].to(DEV), image=torch.rand(B, 3, HW, HW, generator=gen).to(DEV),
                mask=(torch.rand(B, HW, HW, generator=gen) > 0.3).float().to(DEV))


def test_graphed_training_run_equals_the_eager_run_bit_for_bit():
    warm, steps = 2, 4
    # ---- eager: warm-up steps on batch 0, then batches 1..steps
    g_a, optim_a, static_a, fn_a = _setup()
    losses_a = []
    for i in [0] * warm + list(range(1, steps + 1)):
        for k, v in _batch(i).items():
            static_a[k].copy_(v)
        optim_a.zero_grad(set_to_none=True)
        losses_a.append(fn_a().detach().clone())
        optim_a.step()
    # ---- graph: the same weights, batch 0 in the static buffers during warm-up and capture
    g_b, optim_b, static_b, fn_b = _setup()
    for k, v in _batch(0).items():
        static_b[k].copy_(v)
    step = GraphedStep(fn_b, optim_b, static=static_b, warmup=warm)
    _C.launch_counts.clear()
    losses_b = [step(**_batch(i)).detach().clone() for i in range(1, steps + 1)]
    assert not _C.launch_counts, "a replay must not issue C-ABI launches from Python"
    assert step.replays == steps
    for a, b in zip(losses_a[warm:], losses_b):
        assert torch.equal(a, b), (float(a), float(b))
    assert len({float(x) for x in losses_b}) == steps          # every replay saw its own batch
    moved = 0.0
    for (n, pa), (_, pb) in zip(g_a.named_parameters(), g_b.named_parameters()):
        assert torch.equal(pa, pb), n
    torch.manual_seed(0)
    fresh = Graph(adapt_gan_opt(H=HW, W=HW, sample_intvs=N, device=DEV), n_train_images=B).to(DEV)
    for (n, p0), (_, pb) in zip(fresh.named_parameters(), g_b.named_parameters()):
        moved = max(moved, float((p0 - pb).abs().max()))
        if n.startswith("nerf.mlp_feat"):
            assert torch.equal(p0, pb), n                      # the frozen trunk stays put (layers/nerf_static_transient_light.py:34,87)
    assert moved > 1e-3                                        # ... and the heads / latents were trained inside the graph


def test_graphed_step_argument_checks():
    g, optim, static, fn = _setup()
    plain = torch.optim.Adam(g.nerf.mlp_rgb.parameters(), lr=1e-3)
    with pytest.raises(ValueError, match="capturable"):
        GraphedStep(fn, plain, static=static)
    with pytest.raises(ValueError, match="CUDA tensor"):
        GraphedStep(fn, optim, static=dict(x=torch.zeros(2)))
    for k, v in _batch(0).items():
        static[k].copy_(v)
    step = GraphedStep(fn, optim, static=static, warmup=1)
    with pytest.raises(KeyError):
        step(depth=torch.zeros(1, device=DEV))
    with pytest.raises(ValueError, match="fixed shapes"):
        step(coords=torch.zeros(B, P, P + 1, 2, device=DEV))
