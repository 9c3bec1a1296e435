"""Loader for the reference's own Python modules staged under oracle/_ref/py by oracle/build_ref.py.

TEST INFRASTRUCTURE ONLY.  Lets the contract tests drive the reference's REAL `render()`
(gaussian_renderer/__init__.py:20-195) and REAL `GaussianModel` (scene/gaussian_model.py:632-1257) on the GPU box,
once over the reference's rasterizer (its own Python wrapper bound to oracle/_ref/ref_dgr_C.so) and once over the
drop-in modules of this repository.  The three packages the reference imports but this image lacks are stubbed:
`plyfile` (PLY IO only), `FrEIA` (imported at gaussian_model.py:28-29, every use commented out) and - when the
reference's kNN is wanted - `simple_knn._C` is pointed at oracle/_ref/ref_knn_C.so.
"""
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
PY = os.path.join(_HERE, "_ref", "py")
_cache = {}


def available():
    return os.path.exists(os.path.join(PY, "gaussian_renderer", "__init__.py")) and \
        os.path.exists(os.path.join(_HERE, "_ref", "ref_dgr_C.so"))


def _stub(name, **attrs):
    if name not in sys.modules:
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
    return sys.modules[name]


def _install_stubs():
    _stub("plyfile", PlyData=None, PlyElement=None)
    f = _stub("FrEIA")
    f.framework = _stub("FrEIA.framework")
    f.modules = _stub("FrEIA.modules")
    if PY not in sys.path:
        sys.path.insert(0, PY)          # `utils`, `scene` (namespace package), `gaussian_renderer`


def gaussian_model():
    """The reference's scene.gaussian_model module (real classes).  `simple_knn._C` resolves to whatever is
    importable (the drop-in package of this repository when tests run)."""
    _install_stubs()
    import scene.gaussian_model as gm
    return gm


def reference_rasterizer_module():
    """The reference's own diff_gaussian_rasterization Python package, its `_C` bound to ref_dgr_C.so."""
    if "ref_dgr" in _cache:
        return _cache["ref_dgr"]
    from . import ref_driver
    C = ref_driver._load("ref_dgr_C")
    name = "ref_dgr_pkg.diff_gaussian_rasterization"
    _stub("ref_dgr_pkg")
    path = os.path.join(PY, "ref_dgr", "diff_gaussian_rasterization", "__init__.py")
    spec = importlib.util.spec_from_file_location(name, path, submodule_search_locations=[os.path.dirname(path)])
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    sys.modules[name + "._C"] = C          # `from . import _C` (reference __init__.py:14) finds it here
    spec.loader.exec_module(m)
    _cache["ref_dgr"] = m
    return m


def render_fn(rasterizer_module):
    """The reference's render() with `diff_gaussian_rasterization` resolved to `rasterizer_module`
    (a fresh copy of gaussian_renderer per rasterizer: its import binds the names at module level)."""
    key = ("render", id(rasterizer_module))
    if key in _cache:
        return _cache[key]
    gaussian_model()
    saved = sys.modules.get("diff_gaussian_rasterization")
    sys.modules["diff_gaussian_rasterization"] = rasterizer_module
    try:
        path = os.path.join(PY, "gaussian_renderer", "__init__.py")
        spec = importlib.util.spec_from_file_location("gaussian_renderer_%d" % len(_cache), path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
    finally:
        if saved is not None:
            sys.modules["diff_gaussian_rasterization"] = saved
        else:
            del sys.modules["diff_gaussian_rasterization"]
    _cache[key] = m.render
    return m.render
