#!/bin/bash
# Builds liblzb200.so (sm_100a only) next to this script. Used by __graft_entry__.build() and by hand.
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wno-deprecated-declarations -Xcudafe --diag_suppress=1444 -Wno-deprecated-declarations"
mkdir -p build
pids=()
for f in lz_graph lz_kernels lz_api lz_rank; do
  if [ ! -f build/$f.o ] || [ $f.cu -nt build/$f.o ] || [ -n "$(find . -maxdepth 1 -name '*.h' -newer build/$f.o)" ] || [ ../../include/lz.h -nt build/$f.o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f.cu -o build/$f.o &
    pids+=($!)
  fi
done
if [ ! -f build/lz_host.o ] || [ lz_host.cc -nt build/lz_host.o ] || [ -n "$(find . -maxdepth 1 -name '*.h' -newer build/lz_host.o)" ] || [ ../../include/lz.h -nt build/lz_host.o ]; then
  g++ -O2 -std=c++17 -fPIC -c lz_host.cc -o build/lz_host.o &
  pids+=($!)
fi
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait $p; done
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o liblzb200.so build/lz_graph.o build/lz_kernels.o build/lz_api.o build/lz_rank.o build/lz_host.o -lcudart -ldl
echo "built $(pwd)/liblzb200.so"
