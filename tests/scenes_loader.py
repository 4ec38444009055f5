"""The synthetic workload generator (resvg_b200/scenes.py: plain numpy, no device code) for the CPU-only legs.

Importing the ``resvg_b200`` package loads libresvg_b200.so; the reference arm of bench.py must not, so that the driver's
record of loaded native libraries shows oracle/liboracle.so alone.  scenes.py is therefore loaded from its file."""
import importlib.util
import os
import sys

_PATH = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "resvg_b200", "scenes.py")


def load():
    if "resvg_b200" in sys.modules:
        from resvg_b200 import scenes
        return scenes
    if "rb_scenes_standalone" in sys.modules:
        return sys.modules["rb_scenes_standalone"]
    spec = importlib.util.spec_from_file_location("rb_scenes_standalone", _PATH)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["rb_scenes_standalone"] = mod
    spec.loader.exec_module(mod)
    return mod
