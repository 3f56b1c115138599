"""Drop-in entry point: `from IrrLanczos import IrrLanczos` keeps working for the reference's
drivers (Irr3Ddeuteron.py:38); the class is the B200-native one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from lanczos_b200.irregular import IrrLanczos  # noqa: E402,F401
