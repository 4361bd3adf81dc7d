#!/usr/bin/env python
"""Build A/B variants of the library with -D switches (dev tool):
    python tools/ab_variants.py name1:-DX=1,-DY=0 name2:...
-> cerberusnet_b200/build/variants/lib_<name>.so ; run with CERB_LIB_OVERRIDE=<path>."""
import os, shutil, subprocess, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cerberusnet_b200 import build as B

out = os.path.join(B.HERE, "build", "variants")
os.makedirs(out, exist_ok=True)
keep = B.LIB_PATH + ".keep"
shutil.copy2(B.LIB_PATH, keep)
try:
    for spec in sys.argv[1:]:
        name, _, flags = spec.partition(":")
        B.build_library(force=True, extra_flags=[f for f in flags.split(",") if f])
        shutil.copy2(B.LIB_PATH, os.path.join(out, f"lib_{name}.so"))
        print("built", name)
finally:
    shutil.move(keep, B.LIB_PATH)
