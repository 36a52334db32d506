#!/usr/bin/env python3
"""Prints a source file with its local `#include "..."` lines replaced by the files they name (recursively, each
file once), wrapped as a C++ raw string literal: the text NVRTC compiles cannot include from disk.

    python3 tools/flatten_includes.py maray_b200/csrc/device_sem.cuh > maray_b200/csrc/device_sem_text.inc
"""
import os
import re
import sys


def flatten(path, seen, out):
    base = os.path.dirname(path)
    for line in open(path):
        m = re.match(r'\s*#include\s+"([^"]+)"', line)
        if m and os.path.exists(os.path.join(base, m.group(1))):
            inc = os.path.normpath(os.path.join(base, m.group(1)))
            if inc not in seen:
                seen.add(inc)
                flatten(inc, seen, out)
        else:
            out.append(line)


def main():
    out = []
    flatten(os.path.normpath(sys.argv[1]), set(), out)
    text = "".join(out)
    assert ')MRSEM"' not in text
    # MSVC-free toolchain, but keep every literal piece well under 64 KiB anyway: adjacent literals concatenate.
    sys.stdout.write('R"MRSEM(')
    size = 0
    for line in text.splitlines(keepends=True):
        if size + len(line) > 60000:
            sys.stdout.write(')MRSEM"\nR"MRSEM(')
            size = 0
        sys.stdout.write(line)
        size += len(line)
    sys.stdout.write(')MRSEM"\n')


if __name__ == "__main__":
    main()
