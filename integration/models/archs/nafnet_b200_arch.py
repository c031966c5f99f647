"""Drop-in arch file for the reference tree: copy to ``<reference>/models/archs/nafnet_b200_arch.py``.

Exports the sm_100a ``NAFNet`` / ``NAFNetRefFusion`` and re-binds those names inside the stock
``network_nafnet_guided_arch.py`` so that the reference's first-match registry (models/archs/__init__.py:33-46) resolves
``type: NAFNetRefFusion`` (options/train_restoration/002*.yml) here whatever the scan order; the stock classes are kept as
``StockNAFNet`` / ``StockNAFNetRefFusion`` for inspection (to run them load the stock file under another module name).  ``reffusion_n_blocks`` may have 4 entries as in option 002 (the
stock class needs 5 and raises IndexError): the last entry is reused for the middle fusion stage.
"""
import importlib

from textualdegremoval_b200.archs.nafnet_b200_arch import NAFNet, NAFNetRefFusion  # noqa: F401

_NAMES = ("NAFNet", "NAFNetRefFusion")


def _rebind(stem):
    try:
        stock = importlib.import_module(f"{__package__}.{stem}") if __package__ else None
    except ImportError:
        stock = None
    if stock is None:
        return
    for name in _NAMES:
        cur = getattr(stock, name, None)
        if cur is not None and cur is not globals()[name]:
            setattr(stock, "Stock" + name, cur)
            setattr(stock, name, globals()[name])


_rebind("network_nafnet_guided_arch")
