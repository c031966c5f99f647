"""Drop-in arch file for the reference tree: copy to ``<reference>/models/archs/promptir_b200_arch.py``.

Same mechanism as ``restormer_b200_arch.py`` next to it: exports the B200 ``PromptIRRefFusion`` and re-binds the name inside
the stock ``network_promptir_guided_arch`` module, so that ``type: PromptIRRefFusion``
(options/train_restoration/001_promptir_all_in_one_restoration.yml) resolves to the sm_100a implementation whichever
module the registry's ``os.scandir`` reaches first.  Inference only; ``decoder: True`` is the only mode whose forward
runs -- in the reference too (with ``decoder: False`` both raise the same shape error at ``up4_3``).
"""
import importlib

from textualdegremoval_b200.archs.promptir_b200_arch import PromptIRRefFusion  # noqa: F401

_NAMES = ("PromptIRRefFusion",)


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


_rebind("network_promptir_guided_arch")
