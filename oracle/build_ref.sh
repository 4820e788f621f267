#!/usr/bin/env bash
# oracle/build_ref.sh — TEST INFRASTRUCTURE.  Compiles the REFERENCE's own three compute shaders, from where they lie
# under /root/reference, into oracle/_ref/libcloudsky_ref.so (git-ignored; travels to the GPU box with the snapshot).
#
# No reference source is copied into the repo: each shader is streamed through the two edits below straight into g++.
# The ONLY edits made to the shader text (everything else GLSL needs is supplied by oracle/glsl_compat.h as C++):
#   1. the two lines `#[compute]` (Godot's section marker) and `#version 450` are dropped — not C++ preprocessor directives;
#   2. an `out T name` parameter qualifier becomes `T& name` (C++ has no keyword that makes the FOLLOWING type a reference);
#      only sky-lut.glsl / transmittance-lut.glsl have any (get_atmosphere_collision_coefficients, compute_inscattering).
# Flags: -fsingle-precision-constant gives unsuffixed literals GLSL's float type; -ffp-contract=off forbids FMA contraction.
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
ref="${CLOUDSKY_REFERENCE_DIR:-/root/reference}/cloud_sky"
out="$here/_ref"
CXX="${CXX:-g++}"
FLAGS="-O2 -std=c++17 -fPIC -pthread -fsingle-precision-constant -ffp-contract=off -fno-fast-math -Wall -Wno-unused-variable -Wno-unused-but-set-variable -Wno-unused-function"

if [ ! -f "$ref/clouds.glsl" ]; then
    echo "build_ref.sh: $ref/clouds.glsl not found (the reference tree is only mounted in the build container)" >&2
    exit 3
fi
mkdir -p "$out"

unit() {  # $1 = namespace tag, $2 = shader file
    echo '#include "glsl_compat.h"'
    echo "namespace refns_$1 {"
    echo 'GLSL_USING_BUILTINS'
    echo "#line 1 \"$2\""
    sed -e '/^#\[compute\]$/d' -e '/^#version 450$/d' -e 's/\bout \(vec[234]\|float\) /\1\& /g' "$2"
    echo '}'
    echo "#define REF_GLUE_$1"
    echo '#include "ref_glue.inc"'
}

unit clouds "$ref/clouds.glsl" | $CXX $FLAGS -I"$here" -x c++ -c - -o "$out/ref_clouds.o"
unit sky "$ref/sky-lut.glsl" | $CXX $FLAGS -I"$here" -x c++ -c - -o "$out/ref_sky.o"
unit transmittance "$ref/transmittance-lut.glsl" | $CXX $FLAGS -I"$here" -x c++ -c - -o "$out/ref_transmittance.o"
$CXX -shared -pthread -o "$out/libcloudsky_ref.so" "$out/ref_clouds.o" "$out/ref_sky.o" "$out/ref_transmittance.o"
rm -f "$out"/ref_*.o
( cd "$ref" && sha256sum clouds.glsl sky-lut.glsl transmittance-lut.glsl ) > "$out/SOURCES.sha256"
echo "built $out/libcloudsky_ref.so"
