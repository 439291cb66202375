#!/bin/bash
# A/B of kernel build variants on the GPU box: tools/ab_variants.sh <tag> [bench args...] ; libs = wave-fenics_b200/libwavefx*.so
# Prints one line per library: tag, library, ms per apply (median window), fraction of the HBM peak.
tag=$1; shift
for lib in wave-fenics_b200/libwavefx*.so; do
  [ -f "$lib" ] || continue
  out=$(WFX_LIB=$PWD/$lib python bench.py --no-rk4 --no-cpu-baseline --no-affine --windows 5 --steps 20 --sustain-s 0 "$@" 2>&1 | tail -1)
  echo "$tag $(basename $lib) $(echo "$out" | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f ms  frac %.3f  windows %s  parity %s' % (d['ms_per_step'], d['roofline']['frac'], [round(w/d['steps'],4) for w in d['windows_ms']], d['parity']))" 2>&1 | tail -1)"
done
