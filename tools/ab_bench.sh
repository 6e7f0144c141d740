#!/bin/bash
# A/B two builds of libinr_b200.so on ONE box: tools/ab_bench.sh <lib_a.so> <lib_b.so> [bench args...]; alternates a b a b.
A=$1; B=$2; shift 2
PKG=mri_implicit_neural_representations_b200
cp $PKG/libinr_b200.so /tmp/lib_keep.so
for v in $A $B $A $B; do
  cp $v $PKG/libinr_b200.so
  python bench.py "$@" --no-gpu-eager --no-secondary --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step']*1e3,2), 'us  e2e', round(d['e2e']['ms_per_step']*1e3,2), {k: round(v*1e3,1) for k,v in d['roofline']['kernels_ms'].items()})"
done
cp /tmp/lib_keep.so $PKG/libinr_b200.so
