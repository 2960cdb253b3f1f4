#!/bin/bash
# Development build with the attention event trace compiled in (-DPT_ATTN_TRACE): ab/libpt_trace.so, loaded through PT_B200_LIB by
# tools/attn_trace.py (per-warp timeline of CTA 0) and tools/attn_hang_hunt.py (where a dead pipeline is waiting).  The product
# library (build.sh) contains none of it.  Run ./build.sh first: every other object is taken from build/.
set -e
cd "$(dirname "$0")/.."
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
mkdir -p build_tr ab
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC --expt-relaxed-constexpr -DPT_ATTN_TRACE \
  -c prompt_tts_b200/csrc/attention_tcgen05.cu -o build_tr/attention_tcgen05.o
objs=""
for f in build/*.o; do
  if [ "$(basename $f)" = attention_tcgen05.o ]; then objs="$objs build_tr/attention_tcgen05.o"; else objs="$objs $f"; fi
done
$NVCC -shared -o ab/libpt_trace.so $objs -lcudart
echo "built ab/libpt_trace.so"
