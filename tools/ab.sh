#!/bin/bash
# Same-call A/B of two builds of libtipe_rt.so on one B200 box (boxes differ by +-1 %, so a change is judged only inside one
# gpurun call, variants alternating twice).
#   tools/ab.sh "<make EXTRA for variant new, e.g. -DPT_MIN_BLOCKS=3>" "<tune.py arguments, e.g. C2>" ["<second tune.py argument list>" ...]
# Variant `base` is the tree as it is; variant `new` is the same tree compiled with the given EXTRA flags (or, with EXTRA = "-",
# whatever libtipe_rt.so is in tmp_variants/lib_new.so already).  Checksums printed by tune.py must match when the change is
# meant to be exact.
set -e
cd "$(dirname "$0")/.."
extra="$1"; shift
mkdir -p tmp_variants
make -s -C tipe_raytracer_b200/csrc >/dev/null
cp tipe_raytracer_b200/libtipe_rt.so tmp_variants/lib_base.so
if [ "$extra" != "-" ]; then
  touch tipe_raytracer_b200/csrc/*.cu
  make -s -C tipe_raytracer_b200/csrc EXTRA="$extra" >/dev/null
  cp tipe_raytracer_b200/libtipe_rt.so tmp_variants/lib_new.so
  touch tipe_raytracer_b200/csrc/*.cu
  make -s -C tipe_raytracer_b200/csrc >/dev/null
fi
cmd=""
for args in "$@"; do cmd="$cmd python tools/tune.py $args;"; done
/usr/local/graft/bin/gpurun --timeout 900 -- "for r in 1 2; do for v in base new; do cp tmp_variants/lib_\$v.so tipe_raytracer_b200/libtipe_rt.so; echo VARIANT=\$v; $cmd done; done" 2>&1 | grep -E "VARIANT|ms="
cp tmp_variants/lib_base.so tipe_raytracer_b200/libtipe_rt.so
rm -rf tmp_variants
