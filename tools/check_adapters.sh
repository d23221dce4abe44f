#!/bin/bash
# Syntax / type check of every binding in adapters/ against stand-ins of the reference headers (the build image has
# no OpenCV / Eigen, so adapters/ is not compiled by __graft_entry__.build()).  The stand-ins (tools/adapter_stubs/) copy
# names, types, constness, declaration order and access levels of the members the bindings use from
# /root/reference/include/*.h.
cd "$(dirname "$0")/.."
rc=0
for f in adapters/*.cc adapters/*.cpp; do
  if g++ -std=c++14 -fsyntax-only -Wall -Wextra -DMSL_SURFEL_RESIDENT -Itools/adapter_stubs -Iinclude "$f"; then echo "ok  $f"; else echo "FAILED  $f"; rc=1; fi
done
# Second pass, where /root/reference exists: the three bindings whose class headers only need OpenCV / Eigen are checked
# against the REFERENCE'S OWN headers (include/ORBextractor.h, include/SurfelFusion.h, include/PlaneExtractor.h + the peac
# fitter) on top of the stand-in OpenCV / Eigen of oracle/ref_shim_cv/ -- the headers the reference sources themselves are
# compiled against for oracle/_ref/.
REF=${REF:-/root/reference}
if [ -d "$REF/include" ]; then
  for f in adapters/ORBextractor_msl.cc adapters/SurfelFusion_msl.cpp adapters/PlaneExtractor_msl.cpp; do
    if g++ -std=c++14 -fsyntax-only -w -Ioracle/ref_shim_cv -Ioracle -I"$REF/include" -I"$REF" -Iinclude "$f"; then echo "ok  $f (reference headers)"; else echo "FAILED  $f (reference headers)"; rc=1; fi
  done
fi
exit $rc
