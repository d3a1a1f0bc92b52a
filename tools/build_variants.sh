#!/bin/bash
# tools/build_variants.sh NAME:"-DFLAG ..." ...   -- A/B builds of libbean_b200.so that differ in bean_svi.cu's compile flags.
# (VARY = the translation units that get the flags, default bean_svi.cu; several: VARY="a.cu b.cu").  Every other unit is compiled once; results in crispr_bean_b200/variants/libbean_b200_NAME.so (git-ignored, travels
# with the gpurun snapshot).  Select one at run time with BEAN_B200_LIB=<path>.
set -e
cd "$(dirname "$0")/.."
SRC=crispr_bean_b200/csrc
OUT=crispr_bean_b200/variants
VARY=${VARY:-bean_svi.cu}
mkdir -p $OUT/obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr"
objs=""
for f in $SRC/*.cu; do
  b=$(basename $f .cu)
  case " $VARY " in *" $b.cu "*) continue;; esac
  if [ ! -f $OUT/obj/$b.o ] || [ $f -nt $OUT/obj/$b.o ] || [ -n "$(find $SRC include -name '*.cuh' -newer $OUT/obj/$b.o -o -name '*.h' -newer $OUT/obj/$b.o)" ]; then
    nvcc $FLAGS -c $f -o $OUT/obj/$b.o &
  fi
  objs="$objs $OUT/obj/$b.o"
done
wait
for spec in "$@"; do
  name=${spec%%:*}; extra=${spec#*:}
  ( vobjs=""; ok=1
    for vf in $VARY; do
      nvcc $FLAGS $extra -c $SRC/$vf -o $OUT/obj/vary_${name}_${vf%.cu}.o || ok=0
      vobjs="$vobjs $OUT/obj/vary_${name}_${vf%.cu}.o"
    done
    [ $ok = 1 ] && nvcc -shared -gencode arch=compute_100a,code=sm_100a $objs $vobjs -o $OUT/libbean_b200_$name.so && echo built $name ) &
done
wait
ls -la $OUT/*.so
