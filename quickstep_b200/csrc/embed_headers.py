"""Prints {include name, raw string} initialisers for the headers given on the
command line (Makefile -> build/jit_headers.inc, consumed by qs_jit.cu)."""
import os
import sys

for path in sys.argv[1:]:
    text = open(path).read()
    assert ')QSJIT"' not in text
    # split long headers: a raw string literal is limited to 64 KB on some compilers
    print('{"%s", R"QSJIT(%s)QSJIT"},' % (os.path.basename(path), text))
