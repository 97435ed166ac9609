"""B200-native UNet++ hot path. The directory is laid out like the reference checkout (flat modules archs / losses /
metrics / utils / train / val), so `sys.path.insert(0, '<repo>/pytorch_nested-unet_b200'); import archs` is the
drop-in; importing this package does that path setup and re-exports the modules."""
import os as _os
import sys as _sys

_HERE = _os.path.dirname(_os.path.abspath(__file__))
if _HERE not in _sys.path:
    _sys.path.insert(0, _HERE)
