"""CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle.cpp header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this package.  The product package (openvino_tokenizers_b200) never does.
"""
from .oracle import *  # noqa: F401,F403
