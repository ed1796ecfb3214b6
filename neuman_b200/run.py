"""python -m neuman_b200.run <reference_script.py> [args...]  -- runs a reference entry point
(render_360.py, render_test_views.py, ...) with the hot path rebound to the CUDA library."""
import os
import runpy
import sys

from .dropin import install

if __name__ == "__main__":
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = os.path.abspath(sys.argv[1])
    install(os.path.dirname(script))
    sys.argv = sys.argv[1:]
    runpy.run_path(script, run_name="__main__")
