#!/bin/bash
# A/B of library variants on the GPU box:  tools/ab.sh "<lib names>" <knob_sweep args...>   (a lib name X = gym_simpletetris_b200/libst_X.so;
# "product" = the product library)
libs=$1; shift
for l in $libs; do
  echo "== $l"
  if [ "$l" = product ]; then unset ST_B200_LIB; else export ST_B200_LIB=$PWD/gym_simpletetris_b200/libst_$l.so; fi
  python tools/knob_sweep.py "$@" 2>&1 | grep -v Warning
done
