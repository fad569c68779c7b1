#!/bin/bash
# Build count-kernel variants (different NK_ROT_PLAN = which rotate halves run on the FMA pipe)
# into gpurun_out/variants/lib_<name>.so; bench each with NEUROKMER_LIB=... python bench.py
set +e
cd "$(dirname "$0")/.."
OUT=neurokmer_b200/build/variants; mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden"
build() { # name, defines...
  local name=$1; shift
  nvcc $FLAGS "$@" -c neurokmer_b200/csrc/nk_count.cu -o $OUT/nk_count_$name.o 2>&1 | grep -E "error" 
  nvcc $FLAGS "$@" -c neurokmer_b200/csrc/nk_api.cu -o $OUT/nk_api_$name.o 2>&1 | grep -E "error"
  nvcc -shared -o $OUT/lib_$name.so $OUT/nk_count_$name.o neurokmer_b200/build/nk_lif.o neurokmer_b200/build/nk_topn.o \
     neurokmer_b200/build/nk_misc.o $OUT/nk_api_$name.o neurokmer_b200/build/nk_fastx.o -cudart static -lpthread -ldl -lrt 2>/dev/null
}
rm -f $OUT/lib_*.so
build base &
build u2 -DNK_COUNT_UNROLL=2 &
build u8 -DNK_COUNT_UNROLL=8 &
build u16 -DNK_COUNT_UNROLL=16 &
build span2 -DNK_CHUNKS_PER_SPAN=2 &
build span2mb8 -DNK_CHUNKS_PER_SPAN=2 -DNK_COUNT_MINBLOCKS=8 &
build span1mb8 -DNK_CHUNKS_PER_SPAN=1 -DNK_COUNT_MINBLOCKS=8 &
build p27span2mb8 -DNK_ROT_PLAN_ID=27 -DNK_CHUNKS_PER_SPAN=2 -DNK_COUNT_MINBLOCKS=8 &
build p27u8 -DNK_ROT_PLAN_ID=27 -DNK_COUNT_UNROLL=8 &
wait
ls $OUT/*.so
