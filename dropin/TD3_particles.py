"""Drop-in shim: put this directory on PYTHONPATH (before the reference checkout) and the reference's
main.py / evaluate_model.py / q_eval_interface.py / experience_injection.py import the B200 path."""
import os as _os
import sys as _sys

_sys.path.insert(0, _os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))))
from td3_b200.TD3_particles import *  # noqa: F401,F403,E402
from td3_b200 import TD3_particles as _impl  # noqa: E402

globals().update({k: v for k, v in vars(_impl).items() if not k.startswith("__")})
