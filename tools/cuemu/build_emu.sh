#!/bin/bash
# DEVELOPER TOOL: compile the kernels of kmerjs_b200/csrc with g++ against the fiber emulation in this
# directory (ASan + UBSan) into build/emu/libkmerjs_b200_emu.so.  Never part of build(), the tests
# or the bench; see cuda_runtime.h here.
set -e
cd "$(dirname "$0")/../.."
mkdir -p build/emu
SAN=${SAN:--fsanitize=address,undefined -fno-sanitize-recover=undefined}
OPT=${OPT:--O1}
FLAGS="-std=c++17 $OPT -g -fPIC -DKJ_CPU_EMU -I tools/cuemu -Wall -Wno-unused-function -Wno-unknown-pragmas -Wno-unused-variable $SAN"
pids=()
for f in kj_ctx.cu kj_count.cu kj_score.cu kj_dbio.cu kj_synth.cu kj_stats.cpp; do
  g++ $FLAGS -x c++ -c kmerjs_b200/csrc/$f -o build/emu/${f%.*}.o &
  pids+=($!)
done
g++ $FLAGS -c tools/cuemu/emu.cpp -o build/emu/emu.o &
pids+=($!)
for p in "${pids[@]}"; do wait $p; done
g++ -shared $SAN -o build/emu/libkmerjs_b200_emu.so build/emu/*.o -lz
echo build/emu/libkmerjs_b200_emu.so
