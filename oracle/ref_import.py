"""Import the REAL reference with stub modules: the read-only checkout at /root/reference in the build container, or the
unmodified copy that `python -m oracle.install_ref` leaves in baseline/_ref (git-ignored; it travels to the GPU box).

TEST / BENCH INFRASTRUCTURE ONLY (see oracle/texpose_oracle.py header).  Used by `oracle/make_golden.py` to generate
`tests/golden/*.npz`, by `bench.py --impl reference` / its eager-GPU "before" number, and by the integration tests that
patch texpose_b200 into the reference's own Graph (INTEGRATION.md option B).

The reference imports third-party packages that are not installed here (easydict, ipdb,
termcolor, pytorch3d, open3d, kornia, lpips, visdom, matplotlib, imageio, ...).  None of them
is on the render hot path, so each gets a permissive stub; `tools/__init__.py` star-imports the
pytorch3d rasteriser wrapper, so the `tools` package itself is replaced by a path-only stub
(recipe: SURVEY.md section 8c).
"""
from __future__ import annotations

import importlib
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root() -> str:
    cands = [os.environ.get("TEXPOSE_REFERENCE"), "/root/reference", os.path.join(_REPO, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isdir(os.path.join(c, "layers")):
            return c
    return "/root/reference"


REF_ROOT = _find_root()

_STUB_ROOTS = ["ipdb", "termcolor", "pytorch3d", "open3d", "kornia", "lpips", "visdom", "matplotlib",
               "mpl_toolkits", "imageio", "plyfile", "trimesh", "tensorboard"]


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


class _StubFinder:
    """Resolves `import root.any.sub.module` to an empty stub for the roots listed above."""

    @staticmethod
    def find_spec(fullname, path=None, target=None):
        if fullname.split(".")[0] in _STUB_ROOTS:
            from importlib.machinery import ModuleSpec
            return ModuleSpec(fullname, _StubFinder, is_package=True)
        return None

    @staticmethod
    def create_module(spec):
        m = _StubModule(spec.name)
        m.__path__ = []
        return m

    @staticmethod
    def exec_module(module):
        if module.__name__ == "termcolor":
            module.colored = lambda s, *a, **k: s
        if module.__name__ == "ipdb":
            module.set_trace = lambda *a, **k: None


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "layers"))


def _install():
    if getattr(_install, "done", False):
        return
    _install.done = True
    if not os.path.isdir(os.path.join(REF_ROOT, "external")):      # baseline/_ref leaves out external/ (SSIM / LPIPS helpers,
        _STUB_ROOTS.append("external")                             # not on the render path)
    sys.meta_path.append(_StubFinder)
    # easydict: attribute dict -- reuse the product's AttrDict (pure container, no arithmetic)
    from texpose_b200.config import AttrDict
    ed = types.ModuleType("easydict")
    ed.EasyDict = AttrDict
    sys.modules.setdefault("easydict", ed)
    # torch.utils.tensorboard pulls the real tensorboard package
    tb = _StubModule("torch.utils.tensorboard")
    sys.modules.setdefault("torch.utils.tensorboard", tb)
    # `tools` package: path-only stub so `tools.ray_sampler` imports without tools/__init__.py
    tools = types.ModuleType("tools")
    tools.__path__ = [os.path.join(REF_ROOT, "tools")]
    sys.modules["tools"] = tools
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)


def load():
    """Returns a namespace with the reference modules on the hot path."""
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    _install()
    ns = types.SimpleNamespace()
    ns.camera = importlib.import_module("camera")
    ns.ray_sampler = importlib.import_module("tools.ray_sampler")
    ns.patch_sampler = importlib.import_module("tools.patch_sampler")
    ns.nerf_stl = importlib.import_module("layers.nerf_static_transient_light")
    ns.nerf_plain = importlib.import_module("layers.nerf")
    ns.graph_mod = importlib.import_module("model.nerf_adapt_st_gan")
    return ns


def load_surfel():
    """compute_surfelinfo.normal_from_depth / compute_box.* (heavier import chain, all stubbed)."""
    _install()
    cwd = os.getcwd()
    try:
        os.chdir(REF_ROOT)   # `import data` resolves relative split files lazily; keep cwd sane
        box = importlib.import_module("compute_box")
        surfel = importlib.import_module("compute_surfelinfo")
    finally:
        os.chdir(cwd)
    return box, surfel


def load_yaml_opt(name="nerf_lm_adapt_gan", H=480, W=640, device="cpu"):
    """Load options/<name>.yaml with the `_parent_` rule of options.py:60-73 (own loader: the
    reference's `options.set` prompts on stdin)."""
    import yaml
    from texpose_b200.config import AttrDict

    def merge(base, over):
        for k, v in over.items():
            if isinstance(v, dict) and isinstance(base.get(k), dict):
                merge(base[k], v)
            else:
                base[k] = v
        return base

    def read(fname):
        with open(os.path.join(REF_ROOT, fname)) as f:
            d = yaml.safe_load(f)
        parents = d.pop("_parent_", None)
        if parents:
            if isinstance(parents, str):
                parents = [parents]
            for p in parents:
                d = merge(read(p), d)
        return d

    opt = AttrDict(read(f"options/{name}.yaml"))
    opt.device = device
    opt.H, opt.W = H, W
    return opt


class cpu_shim:
    """The reference hard-codes `.cuda()` in three places of the render path (model/nerf_adapt_st_gan.py:600,659-667,708).
    To time its own code on the host cores, `.cuda()` is made the identity for the duration of the block -- an environment
    shim like the import stubs above; no line of the reference changes."""

    def __enter__(self):
        import torch
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda t, *a, **k: t
        return self

    def __exit__(self, *exc):
        import torch
        torch.Tensor.cuda = self._orig
        return False


def build_graph(ns, opt, n_images=4, seed=0):
    """Graph without running Graph.__init__ (it downloads VGG19 and calls .cuda(),
    layers/perceptual_loss.py:28-29): attach only the members `render` reads."""
    import torch
    torch.manual_seed(seed)
    G = ns.graph_mod.Graph
    g = G.__new__(G)
    torch.nn.Module.__init__(g)
    g.nerf = ns.nerf_stl.NeRF(opt)
    g.ray_sampler = ns.ray_sampler.RaySampler(opt)
    g.latent_vars_trans = torch.nn.Embedding(n_images, opt.nerf.N_latent_trans)
    g.latent_vars_light = torch.nn.Embedding(n_images, opt.nerf.N_latent_light)
    return g
