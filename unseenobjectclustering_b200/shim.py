"""Reference-side binding: rebind the reference's own call sites to this package.

The reference inserts its `lib/` at sys.path[0] (tools/_init_paths.py:10-18), so shadowing by
PYTHONPATH does not work; instead import the reference modules first, then call install():

    import _init_paths                      # the reference's own path setup
    import networks, fcn.test_dataset
    from unseenobjectclustering_b200 import shim; shim.install()
    # tools/test_net.py / test_images.py continue unchanged:
    #   network = networks.__dict__['seg_resnet34_8s_embedding'](2, cfg.TRAIN.NUM_UNITS, network_data).cuda()
    #   out_label, out_label_refined = test_sample(sample, network, network_crop)

or run a reference tool under the shim:  python -m unseenobjectclustering_b200.shim tools/test_images.py --args...
"""
import runpy
import sys

from . import networks as _networks
from . import test_dataset as _td
from . import mean_shift as _ms
from . import evaluation as _ev


def install(verbose=False):
    """Patch every already-imported reference module. Returns the list of patched attributes."""
    patched = []
    ref_cfg_mod = sys.modules.get("fcn.config")
    if ref_cfg_mod is not None and hasattr(ref_cfg_mod, "cfg"):
        # what the reference's SEGNET.__init__ / clustering_features read from the global cfg (SEG.py:34-38,
        # test_dataset.py:45) is read from the same object at the same moments (the tools call cfg_from_file() later)
        _networks.configure(ref_cfg_mod.cfg, live=True)
        _td._LIVE_CFG[0] = ref_cfg_mod.cfg
        patched.append("cfg (INPUT, TRAIN.FUSION_TYPE, TRAIN.EMBEDDING_NORMALIZATION, TRAIN.EMBEDDING_METRIC read live)")
    ref_networks = sys.modules.get("networks")
    if ref_networks is not None and hasattr(ref_networks, "__dict__"):
        for name in ("seg_resnet34_8s_embedding", "seg_resnet34_8s_embedding_early"):
            ref_networks.__dict__[name] = getattr(_networks, name)
            patched.append("networks." + name)
    ref_td = sys.modules.get("fcn.test_dataset")
    if ref_td is not None:
        for name in ("clustering_features", "crop_rois", "match_label_crop", "filter_labels_depth", "test_sample"):
            setattr(ref_td, name, getattr(_td, name))
            patched.append("fcn.test_dataset." + name)
    if ref_td is not None and hasattr(ref_td, "multilabel_metrics"):
        ref_td.multilabel_metrics = _ev.multilabel_metrics        # the evaluation tail of test_segnet (:310, :328)
        patched.append("fcn.test_dataset.multilabel_metrics")
    ref_ev = sys.modules.get("utils.evaluation")
    if ref_ev is not None:
        ref_ev.multilabel_metrics = _ev.multilabel_metrics
        patched.append("utils.evaluation.multilabel_metrics")
    ref_ms = sys.modules.get("utils.mean_shift")
    if ref_ms is not None:
        for name in ("mean_shift_smart_init", "select_smart_seeds", "seed_hill_climbing_ball", "connected_components"):
            setattr(ref_ms, name, getattr(_ms, name))
            patched.append("utils.mean_shift." + name)
    if verbose:
        print("unseenobjectclustering_b200.shim patched:", ", ".join(patched))
    return patched


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        raise SystemExit("usage: python -m unseenobjectclustering_b200.shim <reference tool .py> [tool args]")
    tool = argv[0]
    sys.argv = argv
    import os
    sys.path.insert(0, os.path.dirname(os.path.abspath(tool)))
    import _init_paths  # noqa: F401  (the reference's tools/_init_paths.py)
    import networks  # noqa: F401
    import fcn.test_dataset  # noqa: F401
    install(verbose=True)
    runpy.run_path(tool, run_name="__main__")


if __name__ == "__main__":
    main()
