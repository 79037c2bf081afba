#!/bin/bash
# DEVELOPER TOOL: run pytest (or any python args after --py) against the emulated library under ASan.
cd "$(dirname "$0")/../.."
export LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0
export PYTHONPATH="$PWD:$PYTHONPATH"
if [ "$1" == "--py" ]; then shift; exec python -c "import tools.cuemu.plugin" -c pass 2>/dev/null || exec python "$@"; fi
exec python -m pytest -p tools.cuemu.plugin "$@"
