# Drop-in package overlaid on the reference tree: keep the reference's same-named directory on the search path
# (gligen_b200/_overlay.py explains why).
from gligen_b200._overlay import extend as _extend

__path__ = _extend(__path__, __name__)
