"""Host buffers for results: page-locked when a CUDA device is present so the engine's
device->host copies run at full PCIe/C2C speed (torch is used only as the allocator)."""
import numpy as np


def pinned_zeros(shape):
    """float64 zeros; pinned host memory if torch + CUDA are available, else plain NumPy
    (host-only unit tests)."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.zeros(tuple(int(s) for s in shape), dtype=torch.float64,
                               pin_memory=True).numpy()
    except Exception:
        pass
    return np.zeros(shape, np.double)


def pinned_empty(shape):
    """Like pinned_zeros without the fill (for buffers that are overwritten completely)."""
    try:
        import torch
        if torch.cuda.is_available():
            return torch.empty(tuple(int(s) for s in shape), dtype=torch.float64,
                               pin_memory=True).numpy()
    except Exception:
        pass
    return np.empty(shape, np.double)
