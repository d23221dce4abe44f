#!/bin/bash
# Syntax / type check of the ORBmatcher and MapPoint bindings against stand-ins of the reference headers (the build image has
# no OpenCV, so adapters/ is not compiled by __graft_entry__.build()).  The stand-ins copy names, types, constness and
# access levels of the members the bindings use from /root/reference/include/*.h.
set -e
cd "$(dirname "$0")/.."
for f in adapters/ORBmatcher_msl.cc adapters/MapPoint_msl.cc; do
  g++ -std=c++14 -fsyntax-only -Wall -Wextra -Itools/adapter_stubs -Iinclude "$f"
  echo "ok  $f"
done
