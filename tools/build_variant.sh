#!/bin/bash
# Build a variant of libclothgpu.so with extra nvcc flags for A/B measurements on the GPU box:
#   tools/build_variant.sh <name> "<extra flags>"   ->  tools/_build/libclothgpu_<name>.so
# select it with CLOTHGPU_LIB_PATH=tools/_build/libclothgpu_<name>.so (the default library is untouched).
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
src=$root/cloth_sim_self_collision_b200/csrc
out=$root/tools/_build/var_$name
mkdir -p "$out"
flags="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -Xcompiler -fPIC -Xcompiler -O2 -Xcompiler -ffp-contract=off -ccbin /usr/bin/g++"
objs=""
for f in context setup assembly linsolve trisolve ilufactor bvh bvhbuild response elemops api; do
  /usr/local/cuda/bin/nvcc $flags "$@" -c "$src/$f.cu" -o "$out/$f.o" &
  objs="$objs $out/$f.o"
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -ccbin /usr/bin/g++ -o "$root/tools/_build/libclothgpu_$name.so" $objs
echo "$root/tools/_build/libclothgpu_$name.so"
