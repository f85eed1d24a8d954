"""B200-native HyRES residual-codec hot path.

Import name: ``hyres_b200`` (a shim package at the repository root extends its
``__path__`` to this directory, whose hyphenated name is not importable).
"""
import os as _os

# The codec pipeline keeps a stream per image in flight, and with the device coder some of them hold a kernel that runs
# for tens of milliseconds.  Streams that share one of the driver's hardware queues (8 by default) would wait behind it,
# so ask for the maximum before the CUDA context exists (a value the user set is left alone).
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import _lib  # noqa: F401,E402
from .models import (LightWeightCheckerboard, RateDistortionLoss, ResidualJPEGCompression,  # noqa: F401
                     get_scale_table)
from .layers import AttentionBlock, CheckboardMaskedConv2d, MultiScaleRefine, conv1x1, conv3x3  # noqa: F401
from .jpeg import TurboJPEGCompression  # noqa: F401
from .pipeline import HostPipeline  # noqa: F401
from .codec_pipeline import CodecPipeline  # noqa: F401
from . import container, spatial  # noqa: F401
from .export import export_model, load_exported  # noqa: F401
