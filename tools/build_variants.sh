#!/bin/bash
# usage: tools/build_variants.sh name1:"-DFOO=1 -DBAR=2" name2:"..."   (here, before a gpurun call)
# Builds goblin_b200/variants/libgoblin_b200_<name>.so with extra nvcc defines, for A/B sweeps of compile-time
# knobs on the GPU box (bench.py / tests pick one with GOBLIN_B200_LIB=<path>).  Never the shipped library.
set -e
cd "$(dirname "$0")/../goblin_b200/csrc"
make -j8 > /dev/null
mkdir -p ../variants build
for spec in "$@"; do
  name="${spec%%:*}"; defs="${spec#*:}"
  (
    /usr/local/cuda/bin/nvcc $defs -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a --fmad=false \
      -Xcompiler -fPIC -I../../include -I. -Xptxas -v -c device.cu -o build/device_$name.o 2> build/device_$name.ptxas.log
    /usr/local/cuda/bin/nvcc -shared -gencode arch=compute_100a,code=sm_100a -o ../variants/libgoblin_b200_$name.so build/host_math.o build/json_reader.o build/obj_loader.o \
      build/bvh_builder.o build/scene_loader.o build/image_io.o build/image_map.o build/capi_host.o build/device_$name.o \
      -cudart shared -lpthread -lz -ldl 2> /dev/null
    echo "$name: $(grep -A2 'k_extendILi0' build/device_$name.ptxas.log | grep -o 'Used [0-9]* registers' | head -1) $(grep -A1 'k_extendILi0' build/device_$name.ptxas.log | grep -o '[0-9]* bytes spill stores' | head -1)"
  ) &
done
wait
