"""Loaders for the reference's own modules out of oracle/_ref/ (oracle/make_ref.py).  TEST / BASELINE INFRASTRUCTURE.

load_reference_layers()  -> the reference's `layers` module (pygcn/layers.py, unmodified), or None when oracle/_ref/
                            was never made (then the callers fall back to the port, oracle/ref_layer_torch.py).
load_reference_models()  -> (layers, utils, models) with the environment shim of SURVEY.md section 0.7 applied in the
                            HARNESS (a stand-in `matplotlib`, `constants` importable), never in the reference files.
                            `layers_module` replaces the module `models.py` imports its GraphConvolution from
                            (`from layers import GraphConvolution`, models.py:4): pass pygcn_b200.layers for the drop-in.
"""
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")


def _by_path(name, path, register=True):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    if register:
        sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def available(*files):
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in (files or ("layers.py",)))


def load_reference_layers():
    if not available("layers.py"):
        return None
    return _by_path("gcnb_ref_layers", os.path.join(REF_DIR, "layers.py"), register=False)


def load_reference_models(layers_module=None):
    if not available("layers.py", "models.py", "utils.py", "constants.py"):
        return None
    if "matplotlib" not in sys.modules:  # utils.py:14 imports pyplot; absent in this image
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    saved = {k: sys.modules.get(k) for k in ("layers", "utils", "models", "constants")}
    try:
        _by_path("constants", os.path.join(REF_DIR, "constants.py"))
        layers = layers_module if layers_module is not None else _by_path("layers", os.path.join(REF_DIR, "layers.py"))
        sys.modules["layers"] = layers
        utils = _by_path("utils", os.path.join(REF_DIR, "utils.py"))
        models = _by_path("models", os.path.join(REF_DIR, "models.py"))
        return layers, utils, models
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
