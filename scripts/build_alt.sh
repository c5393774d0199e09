#!/bin/bash
# Build a compile-time variant of the library for A/B timing: scripts/build_alt.sh NAME VAR=VALUE ...  ->  diff-dope_b200/diffdope/_lib/alt_NAME.so
# (used through DDOPE_B200_LIB=<path>; never a fallback: the default library is always libddope_b200.so)
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
NAME=$1; shift
TMP=$(mktemp -d)
cp "$ROOT"/diff-dope_b200/csrc/*.cu "$ROOT"/diff-dope_b200/csrc/*.cuh "$ROOT"/diff-dope_b200/csrc/*.h "$ROOT"/diff-dope_b200/csrc/Makefile "$TMP"/
sed -i "s#\"../../include/ddope_b200.h\"#\"$ROOT/include/ddope_b200.h\"#" "$TMP"/*.cu
sed -i "s#../../include/ddope_b200.h#$ROOT/include/ddope_b200.h#; s#OUT := ../diffdope/_lib/libddope_b200.so#OUT := $ROOT/diff-dope_b200/diffdope/_lib/alt_$NAME.so#; s#mkdir -p ../diffdope/_lib#mkdir -p $ROOT/diff-dope_b200/diffdope/_lib#" "$TMP"/Makefile
make -C "$TMP" "$@" 2>&1 | grep -E "error|Error|pixel_kernelILi1ELb0ELb0ELb0ELb0E|raster_kernelILb0E" -A2 | grep -E "error|Error|spill|Used" || true
rm -rf "$TMP"
ls -la "$ROOT"/diff-dope_b200/diffdope/_lib/alt_$NAME.so
