#!/bin/bash
# Builds libcarla_ppo_b200.so for sm_100a (cross-compiles without a GPU).
set -e
cd "$(dirname "$0")"
OUT=../libcarla_ppo_b200.so
SRCS="vae_api.cu tapgemm.cu wgrad.cu elementwise.cu ppo.cu"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
mkdir -p ../build
OBJS=""
for f in $SRCS; do
  o=../build/${f%.cu}.o
  if [ ! -f $o ] || [ $f -nt $o ] || [ common.cuh -nt $o ] || [ ../../include/carla_ppo_b200.h -nt $o ] || ls *.cuh | xargs -I{} test {} -nt $o 2>/dev/null; then
    nvcc $FLAGS ${NVCC_EXTRA} -c $f -o $o &
  fi
  OBJS="$OBJS $o"
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a $OBJS -o $OUT
echo built $OUT
