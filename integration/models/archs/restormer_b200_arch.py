"""Drop-in arch file for the reference tree: copy to ``<reference>/models/archs/restormer_b200_arch.py``.

The reference's registry (models/archs/__init__.py:9-46) imports every ``*_arch.py`` under ``models/archs`` and
``define_network`` instantiates the FIRST scanned module that exposes the requested class name (``os.scandir`` order,
i.e. arbitrary).  The stock ``network_restormer_guided_arch.py`` defines the same names, so besides exporting the B200
classes this file re-binds ``Restormer`` / ``RestormerRefFusion`` inside the stock module: whichever module the scan
reaches first, ``type: Restormer`` / ``type: RestormerRefFusion`` in an option file (options/train_restoration/003*.yml,
004_*.yml, 011-019*.yml) resolves to the sm_100a implementation.  The stock classes are kept as
``network_restormer_guided_arch.StockRestormer`` / ``StockRestormerRefFusion`` for inspection only: their constructors
call ``super(Restormer, self)`` through the re-bound module-level name, so to RUN the stock CPU classes load the stock file
under another module name (as oracle/ref_loader.py does).

Same constructor kwargs, call signature and ``state_dict`` keys as the stock classes (released ``net_g_*.pth`` load with
``strict=True``); CUDA (sm_100a) tensors only.
"""
import importlib

from textualdegremoval_b200.archs.restormer_b200_arch import Restormer, RestormerRefFusion  # noqa: F401

_NAMES = ("Restormer", "RestormerRefFusion")


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


_rebind("network_restormer_guided_arch")
