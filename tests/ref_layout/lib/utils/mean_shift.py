"""Stand-in for the reference's lib/utils/mean_shift.py (names only)."""


def _not_rebound(name):
    def f(*args, **kwargs):
        raise RuntimeError("stand-in utils.mean_shift.%s was called: shim.install() did not rebind it" % name)
    f.__name__ = name
    return f


mean_shift_smart_init = _not_rebound("mean_shift_smart_init")
select_smart_seeds = _not_rebound("select_smart_seeds")
seed_hill_climbing_ball = _not_rebound("seed_hill_climbing_ball")
connected_components = _not_rebound("connected_components")
