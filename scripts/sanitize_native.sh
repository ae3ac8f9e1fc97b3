#!/bin/bash
# Host runtime under sanitizers: builds csrc/native.cpp with ASan+UBSan (or TSan) into a scratch
# directory and runs the evaluator / codec / hashing tests against that build.
#   bash scripts/sanitize_native.sh [address|thread]
set -e
MODE=${1:-address}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$(mktemp -d /tmp/bsnative-$MODE-XXXX)
PYINC=$(python -c "import sysconfig;print(sysconfig.get_paths()['include'])")
PBINC=$(python -c "import pybind11;print(pybind11.get_include())")
EXT=$(python -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
if [ "$MODE" = thread ]; then FLAGS="-fsanitize=thread"; LIB=$(gcc -print-file-name=libtsan.so); TESTS="tests/test_eval.py tests/test_chaos.py tests/test_session.py"
else FLAGS="-fsanitize=address,undefined -fno-sanitize-recover=undefined"; LIB=$(gcc -print-file-name=libasan.so); TESTS="tests/test_eval.py tests/test_sliceio.py tests/test_frame.py tests/test_chaos.py tests/test_sortio.py tests/test_session.py"; fi
g++ -O1 -g -fno-omit-frame-pointer $FLAGS -std=c++17 -shared -fPIC -fvisibility=hidden \
    -I"$PYINC" -I"$PBINC" "$ROOT/csrc/native.cpp" -o "$OUT/_bsnative$EXT"
echo "built $OUT/_bsnative$EXT"
cd "$ROOT"
# the package imports its native module from BIGSLICE_NATIVE_DIR when set
# libstdc++ must be resolvable when the sanitizer runtime initialises (its __cxa_throw
# interceptor looks the real one up at start-up; python itself does not link libstdc++)
BIGSLICE_NATIVE_DIR="$OUT" LD_PRELOAD="$LIB $(gcc -print-file-name=libstdc++.so)" ASAN_OPTIONS=detect_leaks=0 \
UBSAN_OPTIONS=print_stacktrace=1 \
TSAN_OPTIONS="report_signal_unsafe=0 halt_on_error=1" \
  timeout 1500 python -m pytest \
      $TESTS -x -q -p no:cacheprovider 2>&1 | grep -E "AddressSanitizer|ThreadSanitizer|runtime error|SUMMARY|passed|failed" | tail -12
