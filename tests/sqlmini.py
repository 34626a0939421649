"""Re-export of the mini SQL front-end (lives in the package so bench.py can use it too)."""
from heavydb_b200.sqlmini import *  # noqa: F401,F403
from heavydb_b200.sqlmini import parse  # noqa: F401
