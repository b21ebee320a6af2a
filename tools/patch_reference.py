#!/usr/bin/env python
"""Produce a copy of a portableRT checkout with the CUDA (B200) backend wired in.

    python tools/patch_reference.py /path/to/portableRT  OUT_DIR

Writes OUT_DIR/include/portableRT/*.hpp and OUT_DIR/src/{backend,intersect_cpu}.cpp: the
reference's files with the four ADDITIVE arms every compiled-in backend needs (the same places
OptiX/HIP/SYCL/Embree occupy), plus this repo's intersect_cuda.hpp and prt_b200.h.  Nothing is
removed or rewritten; with USE_CUDA undefined the result is token-for-token the reference.

  1. backend.hpp          forward declaration `class CUDABackend;` and a `CUDABackend *` arm in
                          the BackendVar variant, before the unconditional CPUBackend* arm
                          (reference: include/portableRT/backend.hpp:48-71)
  2. src/backend.cpp      `#include intersect_cuda.hpp` and a dynamic_cast arm in to_variant
                          (reference: src/backend.cpp:3-21, 24-49)
  3. portableRT.hpp       conditional include in the umbrella header (portableRT.hpp:5-23)
  4. nearest_hits_impl.hpp  same include list (nearest_hits_impl.hpp:4-22)

The output is a build artefact (tests write it under oracle/_ref/, which is git-ignored): no
reference source is ever committed to this repository.
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

GUARD_INC = '#ifdef USE_CUDA\n#include "intersect_cuda.hpp"\n#endif\n\n'


def insert_before(text, marker, addition, what):
    i = text.find(marker)
    if i < 0:
        raise SystemExit(f"patch_reference: marker for {what} not found: {marker!r}")
    return text[:i] + addition + text[i:]


def main():
    if len(sys.argv) != 3:
        raise SystemExit(__doc__)
    ref, out = sys.argv[1], sys.argv[2]
    inc_in = os.path.join(ref, "include", "portableRT")
    inc_out = os.path.join(out, "include", "portableRT")
    src_out = os.path.join(out, "src")
    os.makedirs(inc_out, exist_ok=True)
    os.makedirs(src_out, exist_ok=True)
    for f in os.listdir(inc_in):
        if f.endswith(".hpp"):
            shutil.copy(os.path.join(inc_in, f), os.path.join(inc_out, f))
    shutil.copy(os.path.join(ref, "src", "intersect_cpu.cpp"), os.path.join(src_out, "intersect_cpu.cpp"))
    shutil.copy(os.path.join(ROOT, "include", "portableRT", "intersect_cuda.hpp"), inc_out)
    shutil.copy(os.path.join(ROOT, "include", "prt_b200.h"), inc_out)

    # 1. backend.hpp
    p = os.path.join(inc_out, "backend.hpp")
    t = open(p).read()
    t = insert_before(t, "class CPUBackend;", "class CUDABackend;\n", "forward declaration")
    t = insert_before(t, "    CPUBackend *>;", "#if defined(USE_CUDA)\n    CUDABackend *,\n#endif\n",
                      "variant arm")
    open(p, "w").write(t)

    # 2. src/backend.cpp
    t = open(os.path.join(ref, "src", "backend.cpp")).read()
    t = insert_before(t, "namespace portableRT {",
                      '#ifdef USE_CUDA\n#include "../include/portableRT/intersect_cuda.hpp"\n#endif\n\n',
                      "backend.cpp include")
    t = insert_before(t, "\tif (auto *q = dynamic_cast<CPUBackend *>(backend))",
                      "#ifdef USE_CUDA\n\tif (auto *q = dynamic_cast<CUDABackend *>(backend))\n"
                      "\t\treturn q;\n#endif\n", "to_variant arm")
    open(os.path.join(src_out, "backend.cpp"), "w").write(t)

    # 3. umbrella header, 4. nearest_hits_impl.hpp
    for name, marker in (("portableRT.hpp", '#include "nearest_hits_impl.hpp"'),
                         ("nearest_hits_impl.hpp", '#include "backend.hpp"')):
        p = os.path.join(inc_out, name)
        t = open(p).read()
        t = insert_before(t, marker, GUARD_INC, name)
        open(p, "w").write(t)
    print(f"patched portableRT tree written to {out}")


if __name__ == "__main__":
    main()
