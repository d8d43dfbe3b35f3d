"""Put this directory on PYTHONPATH to activate the dasr_b200 import overlay in every interpreter started from that
environment (the alternative to the site-packages .pth file written by `python -m dasr_b200.install --pth`)."""
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _root not in sys.path:
    sys.path.append(_root)
try:
    from dasr_b200 import overlay as _overlay
    _overlay.activate()
except Exception as _e:          # never break an interpreter start-up
    sys.stderr.write('dasr_b200 overlay not activated: %r\n' % (_e,))
