"""Drop-in entry point: `from Lanczos import Lanczos` keeps working for the reference's
drivers (3Ddeuteron.py:8,94); the class is the B200-native one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lanczos_b200.regular import Lanczos  # noqa: E402,F401
from lanczos_b200.engine import StencilOperator  # noqa: E402,F401
