from enum import Enum


class BMM(Enum):
    """Packed-weight byte order selector kept from the reference (binary/cuda/bmm.py); on sm_100a all three run the
    same XOR/POPC kernel."""
    BSTC32 = 1
    BTC32 = 2
    ADAPTIVE = 3
