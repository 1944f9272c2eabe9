"""Stand-in for the reference's lib/utils/evaluation.py (name only)."""


def multilabel_metrics(prediction, gt, obj_detect_threshold=0.75):
    raise RuntimeError("stand-in utils.evaluation.multilabel_metrics was called: shim.install() did not rebind it")
