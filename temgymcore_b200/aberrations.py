"""``KrivanekCoeffs`` (reference ``src/temgym_core/aberrations.py:5-31``).

The aberration function itself (``W_krivanek`` / ``grad_W_krivanek``,
aberrations.py:42-108) is evaluated inside the CUDA ray kernel
(``csrc/trace.cu``); the 25 coefficients travel in the model descriptor in
the field order below.
"""
from dataclasses import dataclass, fields


@dataclass
class KrivanekCoeffs:
    C10: float = 0.0
    C12: float = 0.0
    phi12: float = 0.0
    C21: float = 0.0
    phi21: float = 0.0
    C23: float = 0.0
    phi23: float = 0.0
    C30: float = 0.0
    C32: float = 0.0
    phi32: float = 0.0
    C34: float = 0.0
    phi34: float = 0.0
    C41: float = 0.0
    phi41: float = 0.0
    C43: float = 0.0
    phi43: float = 0.0
    C45: float = 0.0
    phi45: float = 0.0
    C50: float = 0.0
    C52: float = 0.0
    phi52: float = 0.0
    C54: float = 0.0
    phi54: float = 0.0
    C56: float = 0.0
    phi56: float = 0.0

    def as_tuple(self):
        return tuple(float(getattr(self, f.name)) for f in fields(self))
