"""Stand-in for the reference's lib/networks/__init__.py: the two factories of the path; calling one that was not rebound fails."""


def seg_resnet34_8s_embedding(num_classes=2, num_units=64, data=None):
    raise RuntimeError("stand-in networks.seg_resnet34_8s_embedding was called: shim.install() did not rebind it")


def seg_resnet34_8s_embedding_early(num_classes=2, num_units=64, data=None):
    raise RuntimeError("stand-in networks.seg_resnet34_8s_embedding_early was called: shim.install() did not rebind it")
