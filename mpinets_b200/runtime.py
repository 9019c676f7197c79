"""Process-wide engine registry: one Engine (C-ABI context) per CUDA device, created lazily."""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .engine import Engine

_ENGINES: Dict[int, Engine] = {}


def get_engine(device: Optional[torch.device] = None, **kwargs) -> Engine:
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    if idx not in _ENGINES:
        _ENGINES[idx] = Engine(device=idx, **kwargs)
    return _ENGINES[idx]


def set_engine(engine: Engine):
    _ENGINES[engine.device.index] = engine
