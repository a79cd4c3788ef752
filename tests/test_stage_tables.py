"""Host logic of the staged tensor-core kernels, on the CPU: the stage list mlp_tc32 derives from an architecture (SURVEY 8:
layers/nerf_static_transient_light.py:15-61) obeys the protocol the kernel relies on, for the yaml's architecture and for others,
and the chain stage list of the plain model pairs every dz slot with the activation that masks it."""
import pytest
import torch

from texpose_b200 import mlp_tc32, mlp_tc_plain
from texpose_b200.layers._mlp import MLPConfig


def _params(n_feat, skip, n_rgb, n_trans, vc=27, n_light=48, n_tl=16):
    lin = lambda o, i: (torch.zeros(o, i), torch.zeros(o))
    feat = [lin(257 if li == n_feat - 1 else 256, 63 if li == 0 else 256 + (63 if li in skip else 0)) for li in range(n_feat)]
    rgb = [lin(256, 256 + vc + 3 + n_light)] + [lin(256, 256) for _ in range(n_rgb - 2)] + [lin(3, 256)]
    trans = [lin(256, 256 + n_tl)] + [lin(256, 256) for _ in range(n_trans - 2)] + [lin(5, 256)]
    cfg = MLPConfig(L_3D=10, L_view=4, skip=tuple(skip), view_dep=True, n_feat=n_feat, n_rgb=n_rgb, n_trans=n_trans,
                    n_latent_light=n_light, n_latent_trans=n_tl)
    return cfg, feat, rgb, trans


@pytest.mark.parametrize("n_feat,skip,n_rgb,n_trans,static_only", [(8, (4,), 4, 4, False), (8, (4,), 4, 4, True), (6, (2,), 3, 2, False),
                                                                    (10, (3, 6), 2, 5, False), (2, (), 2, 2, False)])
def test_stage_list_protocol(n_feat, skip, n_rgb, n_trans, static_only):
    cfg, feat, rgb, trans = _params(n_feat, skip, n_rgb, n_trans)
    slots, stages, biases = mlp_tc32.build_tables(cfg, feat, rgb, trans, static_only=static_only, save=True)
    H, D, RO, TO = mlp_tc32.KIND_HIDDEN, mlp_tc32.KIND_DENSITY, mlp_tc32.KIND_RGB_OUT, mlp_tc32.KIND_TRANS_OUT
    W, RL, PK, EL = mlp_tc32.F_WAIT_READY, mlp_tc32.F_RELOAD, mlp_tc32.F_PARK, mlp_tc32.F_E_LAST
    assert len(slots) == sum((s[0] + s[1]) if s[2] == H else 1 for s in stages)
    assert [s[2] for s in stages].count(D) == 1 and [s[2] for s in stages].count(RO) == 1
    assert [s[2] for s in stages].count(TO) == (0 if static_only else 1)
    assert stages[0][0] == 0 and stages[0][1] == 4 and stages[-1][2] != H
    pending = 0                     # drains whose column groups nobody has consumed yet
    e_seen_last = False
    for i, (a, e, kind, bk, boff, flags, save) in enumerate(stages):
        assert a in (0, 16) and 0 <= e <= 4 and boff % 4 == 0
        if flags & W:
            assert pending == 1 and a == 16
            pending = 0
        if flags & RL:
            assert pending == 0 and stages[i - 1][2] != H and any(s[5] & PK for s in stages[:i])
        if a and not (flags & (W | RL)):
            assert stages[i - 1][2] != H          # the previous stage left the A tile untouched
        if e_seen_last:
            assert e == 0
        if flags & EL:
            e_seen_last = True
        if kind == H:
            assert pending == 0
            pending = 1
            assert save >= 0
        else:
            assert save == -1 and a == 16 and e == 0
    assert pending == 0 and e_seen_last
    assert sorted(s[6] for s in stages if s[6] >= 0) == list(range(sum(1 for s in stages if s[2] == H)))
    assert sum(b.numel() for b in biases) % 4 == 0
    # K steps of a layer cover its input columns exactly once
    for (ptr, ld, row0, rows, col0, cols, kind, _) in slots:
        assert kind in (256, 16) and cols >= 1 and (cols <= 16 if kind == 256 else cols == 256)


def test_plain_chain_stage_list():
    cfg, feat, rgb, _ = _params(8, (4,), 2, 2, n_light=0)
    cfg = MLPConfig(L_3D=10, L_view=4, skip=(4,), view_dep=True, n_feat=8, n_rgb=2, n_trans=0)
    rgb = [(torch.zeros(128, 256 + 27 + 3), torch.zeros(128)), (torch.zeros(3, 128), torch.zeros(3))]
    assert mlp_tc_plain.supported(cfg, feat, rgb)
    assert not mlp_tc_plain.supported(cfg, feat, rgb + rgb[1:])                     # deeper heads: SIMT path
    rows, stages = mlp_tc_plain._bwd_chunks(cfg, feat, mlp_tc_plain._padded_head(rgb))
    nf = 8
    assert len(stages) == nf + 1 and len(rows) == 2 + 8 * nf                         # two thin chunks + eight K=32 chunks per 256-wide layer
    assert stages[0][:4] == [0, 0, 0, 0] and stages[0][4] == nf                      # rgb output layer: thin operand 0, masked by the rgb hidden activation
    assert stages[1][4] == nf - 1 and stages[2][0] == 1 and stages[2][4] == nf - 2   # feature mask; density row enters as thin operand 1
    assert [s[4] for s in stages[3:]] == list(range(nf - 3, -1, -1))                 # h5 ... h0
    assert [s[5] for s in stages] == list(range(nf + 1))                             # dz slots in chain order
    assert all(s[3] == 8 for s in stages[1:]) and all(s[0] == -1 for s in stages[3:])


def test_pretrain_graphs_host_logic():
    """model/nerf_pretrain.py:505-511 against model/nerf_pretrain_env.py:484-485: pose selection; both Graphs build the plain model
    (and a second one under nerf.fine_sampling, :454-455) with the reference's parameter names."""
    import torch
    from texpose_b200.config import AttrDict, env_opt
    from texpose_b200.model import nerf_pretrain, nerf_pretrain_env
    opt = env_opt()
    var = AttrDict(pose=torch.zeros(2, 3, 4), pose_init=torch.ones(2, 3, 4))
    assert nerf_pretrain.Graph.get_pose(opt, var, mode="train") is var.pose_init       # data.pose_source = predicted
    assert nerf_pretrain.Graph.get_pose(opt, var, mode="val") is var.pose
    assert nerf_pretrain_env.Graph.get_pose(opt, var, mode="train") is var.pose
    g = nerf_pretrain_env.Graph(opt)
    names = [n for n, _ in g.named_parameters()]
    assert names[0] == "nerf.mlp_feat.0.weight" and names[-1] == "nerf.mlp_rgb.1.bias" and len(names) == 20
    opt.nerf.fine_sampling = True
    assert any(n.startswith("nerf_fine.") for n, _ in nerf_pretrain.Graph(opt).named_parameters())


def test_descriptor_tables_are_cached_by_content_and_graphed_step_needs_cuda():
    """Host logic of the launch-bound step sizes: `ops.device_table` hands the same tensor back for the same rows (no re-upload, hence
    no stream synchronisation per training step) and a different one for different rows; `GraphedStep` refuses to run without CUDA
    (no CPU path)."""
    import pytest
    import torch
    from texpose_b200 import ops
    from texpose_b200.train_graph import GraphedStep
    rows = [[1, 2, 3], [4, 5, 6]]
    a = ops.device_table(rows, torch.int64, "cpu")
    b = ops.device_table([list(r) for r in rows], torch.int64, "cpu")
    c = ops.device_table([[1, 2, 3], [4, 5, 7]], torch.int64, "cpu")
    d = ops.device_table(rows, torch.int32, "cpu")
    assert a is b and a is not c and a is not d and d.dtype == torch.int32
    assert a.tolist() == rows and c.tolist()[1][2] == 7
    flat = ops.device_table([3, 1, 2], torch.int32, "cpu")
    assert flat.tolist() == [3, 1, 2] and ops.device_table((3, 1, 2), torch.int32, "cpu") is flat
    for i in range(300):                                  # bounded: the oldest entries are dropped, the newest stay
        ops.device_table([i, i + 1], torch.int64, "cpu")
    assert len(ops._TABLES) <= 257
    assert ops.device_table([299, 300], torch.int64, "cpu").tolist() == [299, 300]
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            GraphedStep(lambda: None)
