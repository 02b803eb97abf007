#!/bin/bash
# tools/build_variant.sh NAME [extra nvcc flags]: build rav1e_b200/variants/libb200rdo_NAME.so with
# me_kernels.cu compiled with the extra flags (kernel A/B experiments; load it with B200RDO_LIB).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants rav1e_b200/variants
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ \
  --expt-relaxed-constexpr -Iinclude "$@" -c rav1e_b200/csrc/me_kernels.cu -o build/variants/me_kernels_$name.o
objs=$(ls build/obj/*.o | grep -v me_kernels.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o rav1e_b200/variants/libb200rdo_$name.so \
  build/variants/me_kernels_$name.o $objs -lcudart_static -lpthread -ldl -lrt
echo rav1e_b200/variants/libb200rdo_$name.so
