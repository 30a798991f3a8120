#!/bin/bash
# CPU only: what the built library contains.  SASS mnemonic census of libcf_b200.so (the tcgen05 / TMEM / TMA instructions the
# B200 profiling recipe lists) per kernel, and the resource usage of every kernel (cuobjdump --dump-resource-usage: registers,
# shared memory, local-memory stack = spills).   bash tools/sass_census.sh > profiles/r5_sass_census.txt
LIB=collaborativefilteringusingtensorflow_b200/libcf_b200.so
CU=/usr/local/cuda/bin
echo "== $LIB: $(stat -c %s $LIB) bytes, $($CU/cuobjdump -lelf $LIB | grep -c sm_100a) sm_100a cubins of $($CU/cuobjdump -lelf $LIB | wc -l)"
$CU/cuobjdump -sass $LIB > /tmp/cf_sass.txt
echo
echo "== whole library: instruction counts"
for m in UTCHMMA UTCQMMA UTMALDG UTMASTG UTCBAR UTCCP LDTM STTM SYNCS LDGSTS 'RED.E.ADD.F32' 'ATOMG' 'REDG' ' HMMA' 'LDG.E.128' 'STG.E.128'; do
  printf "  %-16s %6d\n" "$m" "$(grep -c -- "$m" /tmp/cf_sass.txt)"
done
echo
echo "== kernels that issue tcgen05 / TMA instructions (UTCHMMA = tcgen05.mma kind::f16, UTMALDG = cp.async.bulk.tensor, LDTM = tcgen05.ld)"
awk '/Function : /{fn=$3} /UTCHMMA/{a[fn]++} /UTMALDG/{b[fn]++} /LDTM/{c[fn]++} /UTCBAR/{d[fn]++} END{for(f in a) printf "  %-90s UTCHMMA %3d UTMALDG %3d LDTM %3d UTCBAR %3d\n", substr(f,1,90), a[f], b[f], c[f], d[f]}' /tmp/cf_sass.txt | sort
echo
echo "== resource usage per kernel (REG, SHARED static bytes, STACK = local memory / spills)"
$CU/cuobjdump --dump-resource-usage $LIB 2>/dev/null | awk '/Function /{fn=$2; sub(/:$/,"",fn)} /REG:/{print "  " substr(fn,1,100) "  " $0}' | sed 's/ CONSTANT\[[0-9]*\]:[0-9]*//g; s/ TEXTURE:0 SURFACE:0 SAMPLER:0//' | sort -u
