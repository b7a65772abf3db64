#!/bin/bash
# A/B experimental library builds: tests/tools/ab.sh <scene> <spp> lib1.so lib2.so ...
scene=$1; spp=$2; shift 2
for lib in "$@"; do echo "== $lib"; RTB200_LIB=$PWD/$lib python tests/tools/profile_render.py $scene $spp 2>&1 | tail -2; done
