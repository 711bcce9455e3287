"""Host-side helpers mirrored from bitorch_engine/utils (only what the low-bit Linear path uses)."""
import torch


def _probe_int_gradients() -> bool:
    """GreenBit ship a patched torch that lets integer tensors require gradients (bitorch_engine/__init__.py:10-29).
    Stock torch raises; we then keep qweight.requires_grad False and carry the weight gradient through
    `privileged_grad` only (SURVEY.md section 7 "Custom torch dependency")."""
    try:
        torch.nn.Parameter(torch.zeros((1,), dtype=torch.uint8), requires_grad=True)
        return True
    except RuntimeError:
        return False


TORCH_INT_GRADIENTS = _probe_int_gradients()
