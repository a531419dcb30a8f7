"""REFERENCE-ARM INFRASTRUCTURE — loads the reference's unmodified Python glue (baseline/_ref, installed byte for byte
by baseline/install_ref.py) with its three missing imports bound to switchable back ends:

    import vren                -> 'ref'     the reference's own CUDA kernels (oracle/_ref/vren_ref.so, built in place from
                                            /root/reference/models/csrc by oracle/build_ref.py)
                                  'ours'    ngp_b200.vren over libngp_b200.so
    import tinycudann as tcnn  -> 'standin' baseline/tcnn_standin.py (plain torch ops; tiny-cuda-nn is absent offline)
                                  'ours'    ngp_b200.tcnn over libngp_b200.so
    from torch_scatter import segment_csr -> baseline/shims/torch_scatter (dead code path in the reference)

The reference modules bind `vren` / `tcnn` as module globals at import, so both names resolve to proxy modules whose
attribute look-ups forward to the currently selected back end: one process can run the same reference code on either
implementation, which is what the drop-in parity test does (tests/test_reference_glue_gpu.py).
"""
import importlib
import importlib.util
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_DIR = os.path.join(HERE, "_ref")


class _Switch(types.ModuleType):
    """module proxy: attribute look-ups go to the selected back end."""

    def __init__(self, name):
        super().__init__(name)
        self.__dict__["_backend"] = None
        self.__dict__["_backend_name"] = None

    def __getattr__(self, key):
        be = self.__dict__["_backend"]
        if be is None:
            raise AttributeError(f"{self.__name__}: no back end selected (RefGlue.use(...))")
        return getattr(be, key)


class RefGlue:
    """The loaded reference modules: .custom_functions .rendering .networks .losses .metrics, plus .use(vren=, tcnn=)."""

    def __init__(self):
        from . import install_ref
        if install_ref.install() is None or not install_ref.verify():
            raise RuntimeError("baseline/_ref is missing or modified: run `python baseline/install_ref.py` where /root/reference is mounted")
        self.vren, self.tcnn = _Switch("vren"), _Switch("tinycudann")
        shim_dir = os.path.join(HERE, "shims")
        saved = {k: sys.modules.get(k) for k in ("vren", "tinycudann", "torch_scatter")}
        sys.modules["vren"], sys.modules["tinycudann"] = self.vren, self.tcnn
        sys.modules.pop("torch_scatter", None)
        sys.path.insert(0, shim_dir)
        try:
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")          # torch.cuda.amp.custom_fwd deprecation in the reference's imports
                pkg = "refglue_models"
                spec = importlib.util.spec_from_file_location(pkg, os.path.join(REF_DIR, "models", "__init__.py"),
                                                              submodule_search_locations=[os.path.join(REF_DIR, "models")])
                mod = importlib.util.module_from_spec(spec)
                sys.modules[pkg] = mod
                spec.loader.exec_module(mod)
                self.custom_functions = importlib.import_module(pkg + ".custom_functions")
                self.rendering = importlib.import_module(pkg + ".rendering")
                self.networks = importlib.import_module(pkg + ".networks")
                self.rendering_noCUDA = importlib.import_module(pkg + ".rendering_noCUDA")

                def load_top(name):
                    sp = importlib.util.spec_from_file_location("refglue_" + name, os.path.join(REF_DIR, name + ".py"))
                    m = importlib.util.module_from_spec(sp)
                    sp.loader.exec_module(m)
                    return m
                self.losses = load_top("losses")
                self.metrics = load_top("metrics")
        finally:
            sys.path.remove(shim_dir)
            for k, v in saved.items():
                if v is None:
                    sys.modules.pop(k, None)
                else:
                    sys.modules[k] = v

    def use(self, vren="ref", tcnn="standin"):
        if vren == "ref":
            from oracle import build_ref
            be = build_ref.load()
            if be is None:
                raise RuntimeError("oracle/_ref/vren_ref.so not built (python oracle/build_ref.py where /root/reference is mounted)")
        elif vren == "ours":
            be = importlib.import_module("ngp_b200.vren")
        else:
            raise ValueError(vren)
        self.vren.__dict__["_backend"], self.vren.__dict__["_backend_name"] = be, vren
        if tcnn == "standin":
            from . import tcnn_standin as tb
        elif tcnn == "ours":
            tb = importlib.import_module("ngp_b200.tcnn")
        else:
            raise ValueError(tcnn)
        self.tcnn.__dict__["_backend"], self.tcnn.__dict__["_backend_name"] = tb, tcnn
        return self


_GLUE = None


def load(vren="ref", tcnn="standin"):
    """-> the process-wide RefGlue with the requested back ends selected."""
    global _GLUE
    if _GLUE is None:
        _GLUE = RefGlue()
    return _GLUE.use(vren=vren, tcnn=tcnn)
