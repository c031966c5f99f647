"""Drop-in arch file for the reference tree: copy to ``<reference>/models/archs/drsformer_b200_arch.py``.

Exports the B200 ``DRSformerRefFusion`` / ``DRSformer200L_SPA_RefFusion`` and re-binds the names inside the stock
``network_drsformer_guided_arch`` / ``network_drsformer_guided_arch_200L_SPA`` modules (same mechanism as
``restormer_b200_arch.py`` next to it), so that options 007-010 resolve to the sm_100a implementation whichever module the
registry's ``os.scandir`` reaches first.  Inference only.  Note that the stock 200L_SPA file cannot construct its own
network (it calls ``functools.partial`` without importing ``functools``); this shim does not depend on it importing.
"""
import importlib

from textualdegremoval_b200.archs.drsformer_b200_arch import DRSformer200L_SPA_RefFusion, DRSformerRefFusion  # noqa: F401


def _rebind(stem, names):
    try:
        stock = importlib.import_module(f"{__package__}.{stem}") if __package__ else None
    except ImportError:
        stock = None
    if stock is None:
        return
    for name in names:
        cur = getattr(stock, name, None)
        if cur is not None and cur is not globals()[name]:
            setattr(stock, "Stock" + name, cur)
            setattr(stock, name, globals()[name])


_rebind("network_drsformer_guided_arch", ("DRSformerRefFusion",))
_rebind("network_drsformer_guided_arch_200L_SPA", ("DRSformer200L_SPA_RefFusion",))
