"""Directory to put on sys.path: it holds the shadow ``layers`` package (see dropin/layers/__init__.py)."""
import os

PATH = os.path.dirname(os.path.abspath(__file__))
