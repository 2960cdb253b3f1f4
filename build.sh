#!/bin/bash
# Builds libpt_b200.so and libpt_seanet.so (sm_100a only) in-tree.  nvcc cross-compiles without a GPU.
set -e
cd "$(dirname "$0")"
SRC=prompt_tts_b200/csrc
OUT=prompt_tts_b200/libpt_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr"
mkdir -p build
objs=""
pids=""
for f in $SRC/*.cu; do
  o=build/$(basename ${f%.cu}).o
  objs="$objs $o"
  if [ ! -f $o ] || [ $f -nt $o ] || [ $SRC/common.cuh -nt $o ] || [ $SRC/tc_common.cuh -nt $o ] || [ $SRC/gemm_tile_table.inc -nt $o ] || [ include/prompt_tts_b200.h -nt $o ]; then
    $NVCC $FLAGS ${PTXAS_V:+-Xptxas -v} -c $f -o $o &
    pids="$pids $!"
  fi
done
for p in $pids; do wait $p; done
$NVCC -shared -o $OUT $objs -lcudart
echo "built $OUT"
# EnCodec SEANet layers (include/prompt_tts_seanet.h): a separate library, one translation unit
SN=prompt_tts_b200/libpt_seanet.so
so=build/seanet.o
if [ ! -f $so ] || [ $SRC/seanet/seanet.cu -nt $so ] || [ $SRC/seanet/seanet_core.h -nt $so ] || [ include/prompt_tts_seanet.h -nt $so ]; then
  $NVCC $FLAGS -Xcompiler -Wno-unknown-pragmas ${PTXAS_V:+-Xptxas -v} -c $SRC/seanet/seanet.cu -o $so
fi
$NVCC -shared -o $SN $so -lcudart
echo "built $SN"
