"""Drop-in for the render path of model/nerf_pretrain_env.py `Graph` (reference :429-700): the pre-training Graph of
model/nerf_pretrain.py with one difference -- `get_pose` always takes the ground-truth pose (:484-485)."""
from __future__ import annotations

from . import nerf_pretrain


class Graph(nerf_pretrain.Graph):

    @staticmethod
    def get_pose(opt, var, mode=None):
        return var.pose
