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
exit $rc
