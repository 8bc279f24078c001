#!/usr/bin/env python
"""Same flags as the reference's generate.py (see layerskip_b200/cli.py)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from layerskip_b200.cli import main_generate

if __name__ == "__main__":
    main_generate()
