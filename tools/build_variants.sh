#!/usr/bin/env bash
# Builds A/B variants of libsvox2_b200.so with different march-kernel tunables into
# svox2_b200/csrc/build/variants/<name>.so ; select one at run time with SVOX2_B200_LIB=<path>.
# usage: tools/build_variants.sh name1:"-DLB_X=1 -DLB_Y=2" name2:"..."
set -euo pipefail
cd "$(dirname "$0")/../svox2_b200/csrc"
mkdir -p build/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  odir="build/variants/$name"; mkdir -p "$odir"
  for f in abi render march queue_march render_abi misc loss optim exchange nvol step; do
    case "$f" in march|queue_march|optim|loss|exchange|step) extra="$flags";; *) extra="";; esac
    if [ -z "$extra" ] && [ -f "build/$f.o" ]; then cp "build/$f.o" "$odir/$f.o"; continue; fi
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a \
      -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr -Xptxas -v $extra \
      -c "$f.cu" -o "$odir/$f.o" 2> "$odir/$f.ptxas.log"
  done
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o "build/variants/$name.so" "$odir"/*.o
  echo "built build/variants/$name.so: $(grep -A1 'lattice_march_kernelILi2ELb1ELb0ELb0' "$odir/march.ptxas.log" | grep -oE 'Used [0-9]+ registers|[0-9]+ bytes spill stores' | tr '\n' ' ')"
done
