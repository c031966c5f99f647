"""Import the *unmodified* reference modules from /root/reference (this container only).

Used by ``oracle/make_golden.py`` and by ``tests/test_oracle_vs_reference.py`` (skipped
when /root/reference is absent, e.g. on the GPU box).  Never used by the product.

Shims (documented in SURVEY.md section 0.1):
  * package ``__init__`` files are stubbed, because ``models/__init__.py`` eagerly imports
    ``metrics`` -> ``skimage`` which is not installed (B9);
  * ``RestormerRefFusion``: its ``Encoder`` returns 4 levels but ``forward`` indexes
    ``feat[4]..feat[1]`` (B1).  ``masa_enc.forward`` is wrapped to return ``[None] + feats``
    which is the only dimensionally consistent reading.  The reference source is not edited.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("TDR_REFERENCE_ROOT", "/root/reference")
if not os.path.isdir(os.path.join(REF_ROOT, "models", "archs")):
    # the GPU box has no /root/reference: use the byte-for-byte copies vendored by the recipe oracle/build_ref.py
    _vend = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
    if os.path.isdir(os.path.join(_vend, "models", "archs")):
        REF_ROOT = _vend


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "archs"))


def _stub_packages():
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    for pkg, rel in (("models", "models"), ("models.archs", "models/archs")):
        if pkg not in sys.modules or not hasattr(sys.modules[pkg], "__path__") \
                or REF_ROOT not in str(getattr(sys.modules[pkg], "__path__", [""])[0]):
            m = types.ModuleType(pkg)
            m.__path__ = [os.path.join(REF_ROOT, rel)]
            sys.modules[pkg] = m


def load_arch(stem: str):
    """Load /root/reference/models/archs/<stem>.py as a module."""
    _stub_packages()
    name = f"models.archs.{stem}"
    if name in sys.modules and getattr(sys.modules[name], "__file__", "").startswith(REF_ROOT):
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(REF_ROOT, "models", "archs", stem + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def restormer(**kw):
    return load_arch("network_restormer_guided_arch").Restormer(**kw)


def restormer_ref_fusion(**kw):
    net = load_arch("network_restormer_guided_arch").RestormerRefFusion(**kw)
    orig = net.masa_enc.forward
    net.masa_enc.forward = lambda x: [None] + orig(x)   # B1 index shim
    return net


def promptir_ref_fusion(**kw):
    net = load_arch("network_promptir_guided_arch").PromptIRRefFusion(**kw)
    orig = net.masa_enc.forward
    net.masa_enc.forward = lambda x: [None] + orig(x)   # B1 index shim (same Encoder class as the guided Restormer)
    return net


def drsformer_ref_fusion(spa=False, **kw):
    if spa:
        # the shipped 200L_SPA file uses functools.partial without importing functools (NameError in Encoder.__init__,
        # i.e. option 007 cannot construct its network upstream): the missing name is injected, the source is untouched
        import functools
        mod = load_arch("network_drsformer_guided_arch_200L_SPA")
        mod.functools = functools
        net = mod.DRSformer200L_SPA_RefFusion(**kw)
    else:
        net = load_arch("network_drsformer_guided_arch").DRSformerRefFusion(**kw)
    orig = net.masa_enc.forward
    net.masa_enc.forward = lambda x: [None] + orig(x)   # B1 index shim (same Encoder class as the guided Restormer)
    return net


def nafnet(**kw):
    return load_arch("network_nafnet_guided_arch").NAFNet(**kw)


def nafnet_ref_fusion(**kw):
    return load_arch("network_nafnet_guided_arch").NAFNetRefFusion(**kw)
