#!/bin/bash
# build_variant.sh NAME "-DRN_..." : a tuning build of librndense with other compile-time constants
# -> retinanet_tensorflow_b200/lib/variants/librndense_NAME.so  (select it with RNDENSE_LIB=...)
set -e
ROOT=$(cd $(dirname $0)/.. && pwd)
OUT=$ROOT/retinanet_tensorflow_b200/lib/variants
mkdir -p $OUT/$1
cd $ROOT/retinanet_tensorflow_b200/csrc
for f in rn_api rn_assign rn_loss rn_loss_tile rn_decode rn_compact rn_shapes rn_exchange; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
    -I$ROOT/include --expt-relaxed-constexpr $2 -c $f.cu -o $OUT/$1/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/librndense_$1.so $OUT/$1/*.o
rm -rf $OUT/$1
