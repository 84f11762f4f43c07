"""B200-native implementation of the iterative neural-style-transfer hot path of
msmink01/text-based-image-style-transfer (multi_style_transfer/).  See DESIGN.md.

The directory name contains hyphens; import it with
    importlib.import_module("text-based-image-style-transfer_b200")
or through the alias module `nst_b200` at the repository root.
"""
from . import _lib, engine  # noqa: F401
from ._lib import NstError  # noqa: F401
from .engine import Net, Plan, gram_chw, style_mix_chw, style_mix_gram  # noqa: F401
from .multi_style_transfer.run_style_transfer import run_multi_style_transfer, StyleTransferSession  # noqa: F401
