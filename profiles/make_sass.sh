#!/bin/bash
# SASS evidence for EVERY kernel of libssunet_b200.so (run after csrc/build.py):
#   bash profiles/make_sass.sh r02
# writes profiles/<tag>_sass_histograms.txt : per object file and kernel -- instruction count, the tcgen05 / TMEM / TMA / mbarrier
#                                             instruction counts (UTCHMMA, UTCBAR, LDTM, UTMALDG, UTMASTG, UTMAREDG, SYNCS) and the
#                                             twelve most frequent opcodes, plus registers / shared memory from ptxas
#        profiles/<tag>_sass_tc_kernels.sass.gz : the full SASS listing of the tensor-core kernel files (conv_tc_halo, conv_tc)
TAG=${1:-r02}
cd "$(dirname "$0")/.."
OBJ=ssunet-gan_b200/csrc/build
OUT=profiles/${TAG}_sass_histograms.txt
: > $OUT
for o in $OBJ/*.o; do
  f=$(basename $o .o)
  echo "==== $f.cu" >> $OUT
  cuobjdump -sass $o | awk '
    function flush() {
      if (name == "") return;
      printf("%s\n    instructions %d | UTCHMMA %d UTCBAR %d LDTM %d UTMALDG %d UTMASTG %d UTMAREDG %d SYNCS %d | top:", name, n, c["UTCHMMA"], c["UTCBAR"], c["LDTM"], c["UTMALDG"], c["UTMASTG"], c["UTMAREDG"], c["SYNCS"]);
      m = 0; for (k in h) { keys[++m] = k }
      for (i = 1; i <= 12 && i <= m; ++i) { best = i; for (j = i + 1; j <= m; ++j) if (h[keys[j]] > h[keys[best]]) best = j; t = keys[i]; keys[i] = keys[best]; keys[best] = t; printf(" %s:%d", keys[i], h[keys[i]]) }
      printf("\n"); delete h; delete c; delete keys; n = 0
    }
    /Function : /{ flush(); name = $3 }
    /^[ \t]+\/\*[0-9a-f]+\*\//{ op = $2; if (op ~ /^@/) op = $3; sub(/\..*/, "", op); sub(/;$/, "", op); if (op != "") { h[op]++; n++; c[op]++ } }
    END{ flush() }' | cu++filt | cut -c1-400 >> $OUT
  awk '/Compiling entry function/{ match($0, /function .* for/); name = substr($0, RSTART + 10, RLENGTH - 15) }
       /Used [0-9]+ registers/{ sub(/^ptxas info    : /, ""); printf("    ptxas %-90s %s\n", name, $0) }' $OBJ/$f.o.ptxas.log | cu++filt | cut -c1-260 >> $OUT
done
{ cuobjdump -sass $OBJ/conv_tc_halo.o; cuobjdump -sass $OBJ/conv_tc.o; } | cu++filt | gzip -9 > profiles/${TAG}_sass_tc_kernels.sass.gz
ls -la $OUT profiles/${TAG}_sass_tc_kernels.sass.gz
