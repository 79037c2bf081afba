#!/bin/bash
# tools/build_variants.sh name1 "flags1" name2 "flags2" ...  ->  kmerjs_b200/variants/<name>.so
# Kernel variants side by side for tools/gpu_variants.sh: only kj_count.cu (the extraction kernels) is recompiled.
set -e
cd "$(dirname "$0")/.."
python -c "import kmerjs_b200.build as b; b.build()" > /dev/null
mkdir -p kmerjs_b200/variants build/var
rm -f kmerjs_b200/variants/*.so
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  ( nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function $flags \
      -c kmerjs_b200/csrc/kj_count.cu -o build/var/$name.o &&
    nvcc -shared -o kmerjs_b200/variants/$name.so build/var/$name.o build/obj/kj_ctx.o build/obj/kj_score.o build/obj/kj_dbio.o \
      build/obj/kj_synth.o build/obj/kj_stats.o -gencode arch=compute_100a,code=sm_100a -lz && echo "built $name ($flags)" ) &
done
wait
ls -la kmerjs_b200/variants/
