"""Drop-in for the reference's Python/Regular/Hamiltonian.py: `from Hamiltonian import Hamiltonian`
(3Ddeuteron.py:73) resolves to the matrix-free, device-evaluated implementation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lanczos_b200.hamiltonian import Hamiltonian  # noqa: E402,F401
