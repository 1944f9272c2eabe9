"""Test driver (this container only, where /root/reference exists): run an UNMODIFIED tool of the reference tree under
`shim.main()`.  The image lacks three of the reference's imports (easydict, matplotlib / transforms3d bits) and has a PyYAML
whose yaml.load() needs a Loader; oracle/ref_harness.py provides the stand-ins, this script adds the yaml default.
With --cpu-cuda-identity, Tensor.cuda() is a no-op so that a CPU-only box gets as far as the first kernel call of THIS
package, which must then fail loudly (there is no CPU path)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]

import ref_harness as rh  # noqa: E402

if __name__ == "__main__":
    argv = sys.argv[1:]
    identity = "--cpu-cuda-identity" in argv
    argv = [a for a in argv if a != "--cpu-cuda-identity"]
    rh.load()                                       # stand-ins for the missing imports; imports the reference's modules
    import yaml
    _load = yaml.load
    yaml.load = lambda stream, Loader=None: _load(stream, Loader=Loader or yaml.FullLoader)
    from unseenobjectclustering_b200 import shim
    if identity:
        with rh.cpu_cuda_identity():
            shim.main(argv)
    else:
        shim.main(argv)
