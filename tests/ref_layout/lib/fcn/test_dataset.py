"""Stand-in for the reference's lib/fcn/test_dataset.py: the names of the path; calling one that was not rebound fails."""
from fcn.config import cfg  # noqa: F401  (the reference's module imports its cfg too: that is how shim.install() finds it)


def _not_rebound(name):
    def f(*args, **kwargs):
        raise RuntimeError("stand-in fcn.test_dataset.%s was called: shim.install() did not rebind it" % name)
    f.__name__ = name
    return f


clustering_features = _not_rebound("clustering_features")
crop_rois = _not_rebound("crop_rois")
match_label_crop = _not_rebound("match_label_crop")
filter_labels_depth = _not_rebound("filter_labels_depth")
test_sample = _not_rebound("test_sample")
multilabel_metrics = _not_rebound("multilabel_metrics")
