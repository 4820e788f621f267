#!/bin/bash
# Development aid: build launch-shape / knob variants of the fast kernel into build/variants/lib_<name>.so.
# usage: tools/build_variants.sh name1:"-DCS_X=1 -DCS_Y=2" name2:"..."   (then: python tools/shape_sweep.py [--flags N])
# e.g.   tools/build_variants.sh scalar:"-DCS_PACKED_F32=0" rec9:"-DCS_REC_MIN_BLOCKS=9" thr20:"-DCS_DIRECT_THRESHOLD=20"
set -e
cd "$(dirname "$0")/../godot-volumetric-cloud-demo-v2_b200/csrc"
make -s >/dev/null
mkdir -p ../../build/variants
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ftz=true -Xptxas -v $defs -c -o ../../build/variants/clouds_fast_$name.o clouds_fast.cu
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../build/variants/lib_$name.so context.o peer.o sky_resource.o composite.o lut_kernels.o clouds_strict.o noise_gen.o ../../build/variants/clouds_fast_$name.o assets.o host_logic.o
  echo "built build/variants/lib_$name.so ($defs)"
done
