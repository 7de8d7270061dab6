"""Put the repository root on sys.path: the scripts in tools/ import aigv_assessor_b200 / tests / oracle from there."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
