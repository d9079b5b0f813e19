#!/bin/bash
# SASS evidence for the tensor-core kernels: per kernel, the count of tcgen05 (UTCHMMA / UTCBAR), TMEM (LDTM) and TMA
# (UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce-add) instructions, plus registers and shared memory from ptxas.  Run after csrc/build.py.
#   bash profiles/make_sass.sh > profiles/r01_sass_evidence.txt
cd "$(dirname "$0")/.."
OBJ=ssunet-gan_b200/csrc/build
for f in conv_tc_halo conv_tc; do
  cuobjdump -sass $OBJ/$f.o | awk -v file=$f '
    /Function : /{ if (name != "") printf("%-110s UTCHMMA %4d  LDTM %3d  UTMALDG %3d  UTMASTG %3d  UTMAREDG %3d  SYNCS %4d\n", name, m, l, t, s, r, y);
                   name = $3; m = l = t = s = y = r = 0 }
    /UTCHMMA/{m++} /LDTM/{l++} /UTMALDG/{t++} /UTMASTG/{s++} /UTMAREDG/{r++} /SYNCS/{y++}
    END{ printf("%-110s UTCHMMA %4d  LDTM %3d  UTMALDG %3d  UTMASTG %3d  SYNCS %4d\n", name, m, l, t, s, y) }' | cu++filt | cut -c1-320
done
echo
echo "ptxas resource usage of the same kernels:"
for f in conv_tc_halo conv_tc; do
  awk '/Compiling entry function/{ match($0, /function .* for/); name = substr($0, RSTART + 10, RLENGTH - 15) }
       /Used [0-9]+ registers/{ sub(/^ptxas info    : /, ""); printf("%-100s %s\n", name, $0) }' $OBJ/$f.o.ptxas.log | cu++filt | cut -c1-220
done
