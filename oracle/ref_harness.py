"""Reference harness: import the UNMODIFIED NVlabs/UnseenObjectClustering Python code on CPU.

TEST INFRASTRUCTURE ONLY.  Used in the build container (where /root/reference exists) to
 (a) pin the oracle restatement in oracle/uoc_oracle.py against the reference itself, and
 (b) generate the golden fixtures under tests/golden/ (see oracle/make_golden.py).
Nothing in the product path, the -m gpu tests, smoke() or bench.py imports this module; the
GPU box has no /root/reference.

Stub recipe follows SURVEY.md section 8(c): easydict / matplotlib / transforms3d stand-ins,
cfg set by hand (cfg_from_file's yaml.load() has no Loader and fails on PyYAML 6),
EMBEDDING_PRETRAIN=False (no model_zoo download), Tensor.cuda -> identity for CPU runs.
"""
import contextlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("UOC_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "lib", "fcn"))


class _EasyDict(dict):
    """Minimal attribute-access dict (easydict is not installed in this image)."""

    def __init__(self, d=None, **kw):
        super().__init__()
        d = dict(d or {}, **kw)
        for k, v in d.items():
            setattr(self, k, v)

    def __setattr__(self, k, v):
        if isinstance(v, dict) and not isinstance(v, _EasyDict):
            v = _EasyDict(v)
        super().__setitem__(k, v)

    __setitem__ = __setattr__

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)


_loaded = {}


def load(device="cpu"):
    """Import the reference modules; returns a namespace with cfg, mean_shift, test_dataset, networks."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    import torch

    lib = os.path.join(REF_ROOT, "lib")
    if lib not in sys.path:
        sys.path.insert(0, lib)
    if "easydict" not in sys.modules:
        m = types.ModuleType("easydict")
        m.EasyDict = _EasyDict
        sys.modules["easydict"] = m
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "transforms3d", "transforms3d.quaternions",
                 "transforms3d.euler"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = types.ModuleType(name)
    q = sys.modules["transforms3d.quaternions"]
    for fn in ("mat2quat", "quat2mat", "qmult"):
        if not hasattr(q, fn):
            setattr(q, fn, lambda *a, **k: None)
    e = sys.modules["transforms3d.euler"]
    for fn in ("mat2euler", "euler2mat", "euler2quat", "quat2euler"):
        if not hasattr(e, fn):
            setattr(e, fn, lambda *a, **k: None)
    if "skimage" not in sys.modules:
        # lib/utils/evaluation.py:94 imports skimage.morphology.disk inside boundary_overlap; skimage is not installed here:
        # provide the one function with skimage's own definition (footprint of the pixels with x^2 + y^2 <= r^2)
        try:
            import skimage.morphology  # noqa: F401
        except Exception:
            import numpy as _np
            sk = types.ModuleType("skimage")
            mo = types.ModuleType("skimage.morphology")

            def _disk(radius, dtype=_np.uint8):
                L = _np.arange(-radius, radius + 1)
                X, Y = _np.meshgrid(L, L)
                return _np.array((X ** 2 + Y ** 2) <= radius ** 2, dtype=dtype)

            mo.disk = _disk
            sk.morphology = mo
            sys.modules["skimage"] = sk
            sys.modules["skimage.morphology"] = mo
    try:
        import torchvision  # noqa: F401  (networks/SEG.py imports it, never uses it)
    except Exception:
        sys.modules["torchvision"] = types.ModuleType("torchvision")

    from fcn.config import cfg
    cfg.TRAIN.EMBEDDING_METRIC = "cosine"
    cfg.INPUT = "RGBD"
    cfg.TRAIN.FUSION_TYPE = "add"
    cfg.TRAIN.EMBEDDING_PRETRAIN = False
    cfg.TRAIN.EMBEDDING_NORMALIZATION = True
    cfg.TEST.VISUALIZE = False
    cfg.device = torch.device(device)

    import utils.mean_shift as mean_shift
    import utils.mask as mask_utils
    import fcn.test_dataset as test_dataset
    import networks

    import utils.evaluation as evaluation
    _loaded.update(cfg=cfg, mean_shift=mean_shift, test_dataset=test_dataset, networks=networks,
                   mask_utils=mask_utils, evaluation=evaluation)
    return types.SimpleNamespace(**_loaded)


@contextlib.contextmanager
def cpu_cuda_identity():
    """test_sample() calls .cuda() unconditionally (test_dataset.py:235-241,261); make it a no-op on CPU."""
    import torch
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def build_network(num_units=64, state_dict=None, seed=0, name="seg_resnet34_8s_embedding", input_type="RGBD",
                  fusion_type="add", normalize=True):
    """Construct the reference network quietly (update_model prints ~900 lines), eval mode.  The variant is set
    the way the reference's tools do it -- through the global cfg SEGNET.__init__ reads (SEG.py:34-38) -- and
    restored afterwards."""
    import torch
    ref = load()
    torch.manual_seed(seed)
    saved = (ref.cfg.INPUT, ref.cfg.TRAIN.FUSION_TYPE, ref.cfg.TRAIN.EMBEDDING_NORMALIZATION)
    ref.cfg.INPUT, ref.cfg.TRAIN.FUSION_TYPE, ref.cfg.TRAIN.EMBEDDING_NORMALIZATION = input_type, fusion_type, normalize
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            net = ref.networks.__dict__[name](2, num_units, state_dict)
    finally:
        ref.cfg.INPUT, ref.cfg.TRAIN.FUSION_TYPE, ref.cfg.TRAIN.EMBEDDING_NORMALIZATION = saved
    return net.eval()


def load_tool(name="test_images"):
    """Import an UNMODIFIED script of the reference's tools/ directory as a module (its __main__ block does not run):
    gives access to read_sample / compute_xyz of tools/test_images.py:96-135."""
    import importlib.util
    load()
    tools = os.path.join(REF_ROOT, "tools")
    if tools not in sys.path:
        sys.path.insert(0, tools)          # for `import _init_paths`
    key = "uoc_ref_tool_" + name
    if key in sys.modules:
        return sys.modules[key]
    spec = importlib.util.spec_from_file_location(key, os.path.join(tools, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    with contextlib.redirect_stdout(io.StringIO()):
        spec.loader.exec_module(mod)
    sys.modules[key] = mod
    return mod

