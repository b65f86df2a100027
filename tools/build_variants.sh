#!/bin/bash
# Build variants of libosph_b200.so that differ in compile-time knobs (-D...) of common.cuh / pair.cu / sph_math.cuh.
#   tools/build_variants.sh name1 "-DOSPH_PAIR_THREADS=128 -DPAIR_CAP=512 -DPAIR_MINB64=4 -DPAIR_MINB32=6"  name2 "..."
# -> offshore-sph_b200/lib/variants/lib_<name>.so; compare with tools/bench_variants.sh name1 name2 ... on the GPU box
# (the .so files travel with gpurun).  Knobs worth measuring are listed in DESIGN.md section 6.
set -e
cd "$(dirname "$0")/../offshore-sph_b200/csrc"
make -s
mkdir -p ../lib/variants
ARCH="-gencode arch=compute_100a,code=sm_100a"
while [ $# -ge 2 ]; do
  name=$1; defs=$2; shift 2
  tmp=$(mktemp -d)
  for f in $(ls *.cu | sed 's/\.cu$//' | grep -v '^pair$'); do
    nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC $defs -c $f.cu -o $tmp/$f.o &
  done
  nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC --use_fast_math $defs -Xptxas -v -c pair.cu -o $tmp/pair.o 2> $tmp/pair.log
  wait
  grep -A2 "k_pairIdLi0ELb1\|k_pairIfLi0ELb0" $tmp/pair.log | grep "Function\|Used\|spill" | sed "s/ptxas info    : //" | paste - - - | sed "s/^/[$name] /"
  nvcc -shared $ARCH -o ../lib/variants/lib_$name.so $tmp/*.o -lcudart -ldl
  python -c "import ctypes,sys; ctypes.CDLL(sys.argv[1])" ../lib/variants/lib_$name.so      # every symbol resolves
  rm -rf $tmp
  echo "built lib/variants/lib_$name.so ($defs)"
done
