#!/usr/bin/env python3
"""embed_src.py <in> <out> <name>: wrap a source file into a C++ raw string constant (the micro-op header that bt_jit.cu hands to NVRTC)."""
import sys
src = open(sys.argv[1]).read()
assert ')BTSRC"' not in src
open(sys.argv[2], "w").write(f'// generated from {sys.argv[1]} by tools/embed_src.py\nstatic const char {sys.argv[3]}[] = R"BTSRC(\n{src})BTSRC";\n')
