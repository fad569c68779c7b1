#!/bin/bash
# Build count-kernel variants into neurokmer_b200/build/variants/lib_<name>.so.  Compile-time knobs:
# NK_ROT_PLAN_ID, NK_COUNT_UNROLL, NK_CHUNKS_PER_SPAN, NK_COUNT_THREADS, NK_COUNT_MINBLOCKS, NK_EXP_NORED, NK_EXP_ADDR32, NK_EXP_REDVAL.
# Bench each with tools/bench_variants.sh (NEUROKMER_LIB=... python bench.py) under gpurun.
# Usage: tools/variants.sh "name1:-DFOO=1 -DBAR=2" "name2:..."   (run `python -m neurokmer_b200.build` first)
set +e
cd "$(dirname "$0")/.."
OUT=neurokmer_b200/build/variants; mkdir -p $OUT
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden"
build() { # name[@alternative nk_count.cu], defines...
  local name=$1; shift
  local src=neurokmer_b200/csrc/nk_count.cu
  if [[ $name == *@* ]]; then src=${name#*@}; name=${name%%@*}; cp $src neurokmer_b200/csrc/.nk_count_$name.cu; src=neurokmer_b200/csrc/.nk_count_$name.cu; fi
  nvcc $FLAGS "$@" -c $src -o $OUT/nk_count_$name.o 2>&1 | grep -E " error"
  nvcc $FLAGS "$@" -c neurokmer_b200/csrc/nk_api.cu -o $OUT/nk_api_$name.o 2>&1 | grep -E " error"
  nvcc -shared -o $OUT/lib_$name.so $OUT/nk_count_$name.o $OUT/nk_api_$name.o neurokmer_b200/build/nk_lif.o neurokmer_b200/build/nk_topn.o \
     neurokmer_b200/build/nk_misc.o neurokmer_b200/build/nk_post.o neurokmer_b200/build/nk_exact.o neurokmer_b200/build/nk_fastx.o \
     neurokmer_b200/build/nk_pack.o neurokmer_b200/build/nk_decomp.o neurokmer_b200/build/nk_multi.o neurokmer_b200/build/nk_parse.o \
     neurokmer_b200/build/nk_ingest.o \
     -cudart static -lpthread -ldl -lrt -lz 2>&1 | grep -v deprecated
  python -c "import ctypes; ctypes.CDLL('$OUT/lib_$name.so')" || echo "lib_$name.so does not load"
}
rm -f $OUT/lib_*.so
build base &
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  build $name $defs &
done
wait
rm -f neurokmer_b200/csrc/.nk_count_*.cu
ls $OUT/*.so
