"""TEST INFRASTRUCTURE — not product code.

Loads the *unmodified* reference (selflein/MG-GAN, mounted read-only at
/root/reference in the build container) so that `oracle/make_golden.py` can
execute it and freeze golden vectors, and so that the CPU test-suite can
cross-check `oracle/mggan_oracle.py` against the live reference when it is
present.  The reference needs three third-party modules that are not in this
image (`test_tube`, `matplotlib`, `shapely`); only their import-time surface
is stubbed here.  Nothing under `mg-gan_b200/` imports this
file, and `/root/reference` does not exist on the GPU box.

The reference package is also called `mggan`; to keep it from colliding with
the product package of the same name it is imported into a private module
namespace by temporarily swapping `sys.modules` / `sys.path`.
"""
import argparse
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MGGAN_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "mggan"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


class _HyperOptArgumentParser(argparse.ArgumentParser):
    # test_tube API used at mggan/model/config.py:5,27,82-133
    def __init__(self, strategy=None, **kw):
        super().__init__(**kw)

    def opt_list(self, *a, options=None, tunable=False, **kw):
        return self.add_argument(*a, **kw)


class _Experiment:
    # used at mggan/model/train.py:678-690, mggan/abstract_train.py:27,36,194,201,273
    def __init__(self, save_dir=None, name="x", debug=False, version=0, **kw):
        self.save_dir, self.name, self.version = str(save_dir), name, version

    def get_data_path(self, name, version):
        p = os.path.join(self.save_dir, name, f"version_{version}")
        os.makedirs(p, exist_ok=True)
        return p

    def log(self, *a, **k):
        pass

    def save(self):
        pass

    def argparse(self, a):
        pass


_LOADED = None


def load_reference():
    """Return a namespace with the reference's modules (imported once).

    The product package `mggan` (if already imported) is moved out of the way
    while the reference imports, then restored; the reference modules stay
    reachable only through the returned namespace.
    """
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True

    saved = {k: v for k, v in sys.modules.items() if k == "mggan" or k.startswith("mggan.")}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path.insert(0, REFERENCE_ROOT)

    stubs = ["test_tube", "matplotlib", "matplotlib.pyplot", "matplotlib.patheffects",
             "matplotlib.patches", "shapely", "shapely.geometry", "shapely.ops"]
    saved_stubs = {k: sys.modules.get(k) for k in stubs}
    _stub("test_tube", HyperOptArgumentParser=_HyperOptArgumentParser, Experiment=_Experiment)
    mpl = _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("matplotlib.patheffects")
    mpl.patches = _stub("matplotlib.patches")
    _stub("shapely")
    _stub("shapely.geometry", Polygon=object, MultiPolygon=object)
    _stub("shapely.ops", unary_union=None)

    try:
        import torch
        import numpy as np
        rng_t = torch.random.get_rng_state()
        rng_n = np.random.get_state()
        ns = types.SimpleNamespace()
        ns.utils = importlib.import_module("mggan.utils")
        ns.common_modules = importlib.import_module("mggan.model.modules.common_modules")
        ns.social = importlib.import_module("mggan.model.modules.social")
        ns.cnn = importlib.import_module("mggan.model.modules.cnn")
        ns.standard = importlib.import_module("mggan.model.modules.standard")
        ns.discriminators = importlib.import_module("mggan.model.modules.discriminators")
        ns.model_factory = importlib.import_module("mggan.model.model_factory")
        ns.config = importlib.import_module("mggan.model.config")
        ns.metrics = importlib.import_module("mggan.metrics")
        ns.manifold = importlib.import_module("mggan.manifold")
        ns.abstract_train = importlib.import_module("mggan.abstract_train")  # re-seeds RNGs at import
        ns.train = importlib.import_module("mggan.model.train")
        ns.Experiment = _Experiment
        # abstract_train.py:14-15 reseeds the global RNGs as an import side effect; undo it.
        torch.random.set_rng_state(rng_t)
        np.random.set_state(rng_n)
    finally:
        ref_mods = {k: v for k, v in sys.modules.items() if k == "mggan" or k.startswith("mggan.")}
        for k in ref_mods:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = saved_path
        for k, v in saved_stubs.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _LOADED = ns
    return ns


def load_reference_module(name):
    """Import one more module of the reference (e.g. "mggan.data_utils.BaseTrajectories", "mggan.evaluation") into a
    private namespace with the same sys.modules / sys.path swap and import-time stubs as `load_reference`."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    sys.dont_write_bytecode = True
    saved = {k: v for k, v in sys.modules.items() if k == "mggan" or k.startswith("mggan.")}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path.insert(0, REFERENCE_ROOT)
    stubs = ["test_tube", "matplotlib", "matplotlib.pyplot", "matplotlib.patheffects",
             "matplotlib.patches", "shapely", "shapely.geometry", "shapely.ops", "seaborn"]
    saved_stubs = {k: sys.modules.get(k) for k in stubs}
    _stub("test_tube", HyperOptArgumentParser=_HyperOptArgumentParser, Experiment=_Experiment)
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.patheffects = _stub("matplotlib.patheffects")
    mpl.patches = _stub("matplotlib.patches", Circle=object)
    _stub("shapely")
    _stub("shapely.geometry", Polygon=object, MultiPolygon=object, Point=object)
    _stub("shapely.ops", unary_union=None, cascaded_union=None)
    _stub("seaborn")
    try:
        import torch
        import numpy as np
        rng_t = torch.random.get_rng_state()
        rng_n = np.random.get_state()
        mod = importlib.import_module(name)
        torch.random.set_rng_state(rng_t)
        np.random.set_state(rng_n)
    finally:
        for k in [k for k in sys.modules if k == "mggan" or k.startswith("mggan.")]:
            del sys.modules[k]
        sys.modules.update(saved)
        sys.path[:] = saved_path
        for k, v in saved_stubs.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    return mod
